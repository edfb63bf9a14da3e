// tests/sched_unit.cpp -- CPU stress of csrc/vmig_sched.h (the writers' per-file queue of a lane): compiled and run by
// tests/test_host.py::test_writer_queue_invariants.  4 producers push 200 000 tasks over 37 keys while 12 workers drain;
// checked: never two workers on one key, tasks of a key come out in push order, every task exactly once, pop() ends
// after close().
#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>
#include "vmig_sched.h"

struct Task { uint32_t key; uint64_t seq; };

int main()
{
    const uint32_t kKeys = 37, kProducers = 4, kWorkers = 12; const uint64_t kPerProducer = 50000;
    vmig::KeyedQueue<Task> q; q.init(kKeys);
    std::vector<std::atomic<int>> on_key(kKeys); std::vector<std::atomic<uint64_t>> next_seq(kKeys), last_seen(kKeys);
    for (auto& a : on_key) a = 0;
    for (auto& a : next_seq) a = 0;
    for (auto& a : last_seen) a = 0;
    std::atomic<uint64_t> consumed{0}; std::atomic<int> bad{0};
    std::mutex key_mu[37];                      // producers of one key take its seq under a lock so push order == seq order
    std::vector<std::thread> th;
    for (uint32_t w = 0; w < kWorkers; w++)
        th.emplace_back([&] {
            uint32_t r; Task t;
            while (q.pop(&r, &t)) {
                if (on_key[r].fetch_add(1) != 0) bad.store(1);                 // a second worker on the same key
                if (t.key != r) bad.store(2);
                if (t.seq != last_seen[r].load() + 1) bad.store(3);            // out of push order (or a duplicate / a hole)
                last_seen[r].store(t.seq);
                for (volatile int spin = 0; spin < (int)(t.seq % 64); spin++) { }
                on_key[r].fetch_sub(1);
                consumed++;
                q.done(r);
            }
        });
    std::vector<std::thread> prod;
    for (uint32_t p = 0; p < kProducers; p++)
        prod.emplace_back([&, p] {
            uint64_t x = 88172645463325252ull + p;
            for (uint64_t i = 0; i < kPerProducer; i++) {
                x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                const uint32_t k = (uint32_t)(x % kKeys);
                std::lock_guard<std::mutex> lk(key_mu[k]);
                q.push(k, Task{k, ++next_seq[k]});
            }
        });
    for (auto& t : prod) t.join();
    while (consumed.load() < kProducers * kPerProducer && !bad.load()) std::this_thread::yield();
    q.close();
    for (auto& t : th) t.join();
    uint64_t total = 0;
    for (auto& a : last_seen) total += a.load();
    if (bad.load() || consumed.load() != kProducers * kPerProducer || total != kProducers * kPerProducer) {
        printf("sched unit FAILED: bad=%d consumed=%llu total=%llu\n", bad.load(), (unsigned long long)consumed.load(), (unsigned long long)total);
        return 1;
    }
    printf("sched unit ok: %llu tasks over %u keys, %u workers\n", (unsigned long long)consumed.load(), kKeys, kWorkers);
    return 0;
}
