// vmig_sched.h -- the writers' queue of a lane (internal; header-only so that tests/sched_unit.cpp can hammer it on CPU).
//
// A destination file takes writes from one thread at a time anyway (inode lock), so two writers on one file only queue up
// behind each other: tasks are queued PER KEY (file) and any idle worker takes the next key that has work and is not being
// worked on right now.  Guarantees: at most one worker per key at any time; tasks of one key are handed out in push order;
// every pushed task is handed out exactly once; pop() returns false only after close() when nothing is ready.
#pragma once
#include <stdint.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <vector>

namespace vmig {

template <class Task>
class KeyedQueue {
    std::mutex mu; std::condition_variable cv; bool closed = false;
    std::vector<std::deque<Task>> fq; std::vector<uint8_t> busy, queued; std::deque<uint32_t> ready;
public:
    void init(size_t n_keys) { fq.resize(n_keys); busy.assign(n_keys, 0); queued.assign(n_keys, 0); }
    void push(uint32_t r, const Task& t) {
        { std::lock_guard<std::mutex> lk(mu); fq[r].push_back(t); if (!busy[r] && !queued[r]) { ready.push_back(r); queued[r] = 1; } }
        cv.notify_one();
    }
    // Blocks until some key has work and no worker; marks that key busy until done(key).
    bool pop(uint32_t* r, Task* t) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return closed || !ready.empty(); });
        if (ready.empty()) return false;
        *r = ready.front(); ready.pop_front(); queued[*r] = 0; busy[*r] = 1;
        *t = fq[*r].front(); fq[*r].pop_front();
        return true;
    }
    void done(uint32_t r) {
        bool more = false;
        { std::lock_guard<std::mutex> lk(mu); busy[r] = 0; if (!fq[r].empty() && !queued[r]) { ready.push_back(r); queued[r] = 1; more = true; } }
        if (more) cv.notify_one();
    }
    void close() { { std::lock_guard<std::mutex> lk(mu); closed = true; } cv.notify_all(); }
};

}  // namespace vmig
