/* tests/c_abi_smoke.c -- the C ABI used from plain C, the way the cgo shim uses it: structs by
 * value from C, nullable tables, errors read back on the calling thread.  Compiled (gcc -std=c99 -Werror) and run by
 * tests/test_host.py::test_c_abi_from_plain_c (no GPU: the data-path calls must fail with VMIG_ENOGPU before touching
 * the destination) and by tests/test_gpu.py::test_c_abi_data_path_from_plain_c (GPU: they must succeed).
 *   usage: c_abi_smoke SRC DST MOVED_DST TABLE
 * Line "layout ..." prints sizeof/offsetof of every ABI struct; the Python tests compare it with their ctypes mirror
 * (and a Go maintainer can compare it with unsafe.Sizeof(C.vmig_opts{})). */
#include <pthread.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "vmig.h"

static void* other_thread(void* out)
{
    /* vmig_last_error is thread-local: a failure on the main thread must not show up here */
    strncpy((char*)out, vmig_last_error(), 255);
    return NULL;
}

int main(int argc, char** argv)
{
    if (argc < 5) return 2;
    const char *src = argv[1], *dst = argv[2], *moved = argv[3], *table = argv[4];
    printf("layout abi=%d opts=%zu:%zu,%zu,%zu,%zu,%zu,%zu stats=%zu:%zu,%zu,%zu,%zu,%zu,%zu,%zu,%zu,%zu,%zu tinfo=%zu:%zu,%zu,%zu\n",
           VMIG_ABI_VERSION, sizeof(vmig_opts), offsetof(vmig_opts, gpu_mask), offsetof(vmig_opts, block_bytes),
           offsetof(vmig_opts, streams_per_gpu), offsetof(vmig_opts, flags), offsetof(vmig_opts, io_threads),
           offsetof(vmig_opts, lanes_per_gpu),
           sizeof(vmig_stats), offsetof(vmig_stats, bytes_total), offsetof(vmig_stats, bytes_d2h), offsetof(vmig_stats, blocks_skipped),
           offsetof(vmig_stats, kernel_launches), offsetof(vmig_stats, ns_total), offsetof(vmig_stats, ns_table),
           offsetof(vmig_stats, ms_kernel), offsetof(vmig_stats, gpus_used), offsetof(vmig_stats, lanes_used),
           offsetof(vmig_stats, pruned),
           sizeof(vmig_table_info), offsetof(vmig_table_info, algo), offsetof(vmig_table_info, n_files),
           offsetof(vmig_table_info, bytes_total));

    int64_t v = 0;
    if (vmig_to_bytes("20GB", &v) != VMIG_OK || v != 21474836480LL) { printf("to_bytes\n"); return 1; }
    if (vmig_to_bytes("1XB", &v) != VMIG_EINVAL) { printf("to_bytes bad unit\n"); return 1; }
    char seen[256]; memset(seen, 0, sizeof seen);
    pthread_t th;
    if (pthread_create(&th, NULL, other_thread, seen) != 0) return 1;
    pthread_join(th, NULL);
    if (seen[0] != 0 || strstr(vmig_last_error(), "XB") == NULL) { printf("last_error is not per thread: main='%s' other='%s'\n", vmig_last_error(), seen); return 1; }

    vmig_stats ms; memset(&ms, 0xAB, sizeof ms);
    if (vmig_manifest(src, 0, 0, NULL, &ms) != VMIG_OK) { printf("manifest: %s\n", vmig_last_error()); return 1; }
    int64_t bytes = 0; uint64_t nf = 0;
    if (vmig_dir_size(src, &bytes, &nf) != VMIG_OK) { printf("dir_size: %s\n", vmig_last_error()); return 1; }
    const int64_t src_dir_size = bytes;          /* utils.DirSize counts every non-directory entry (symlinks too) */
    printf("version=%s files=%llu bytes=%llu dir_size=%lld\n", vmig_version(), (unsigned long long)ms.files,
           (unsigned long long)ms.bytes_total, (long long)bytes);

    /* utils.CopyDir with the tables: opts and stats are plain C structs passed by address, like cgo's C.vmig_opts */
    vmig_opts o; memset(&o, 0, sizeof o);
    o.flags = VMIG_F_VERIFY; o.lanes_per_gpu = 2;
    vmig_stats st; memset(&st, 0xCD, sizeof st);
    int rc = vmig_migrate_tree(src, dst, NULL, table, &o, &st);
    if (rc == VMIG_ENOGPU) { printf("copy refused: %s | %s\n", vmig_strerror(rc), vmig_last_error()); return 0; }
    if (rc != VMIG_OK) { printf("copy failed rc=%d: %s\n", rc, vmig_last_error()); return 1; }
    if (st.bytes_total != ms.bytes_total || st.files != ms.files || st.blocks_total != ms.blocks_total || st.bytes_written == 0 || st.bytes_written > st.bytes_total) {   /* hard-linked paths are written once */
        printf("stats disagree with the manifest pass\n"); return 1;
    }
    vmig_table_info ti;
    if (vmig_table_info_read(table, &ti) != VMIG_OK || ti.n_blocks != st.blocks_total || ti.bytes_total != st.bytes_total) { printf("table info\n"); return 1; }
    printf("copy ok bytes=%llu blocks=%llu lanes=%u launches=%llu\n", (unsigned long long)st.bytes_total, (unsigned long long)st.blocks_total,
           st.lanes_used, (unsigned long long)st.kernel_launches);

    /* second pass against the table just written: nothing changed, so only hard-linked files (never patched in
     * place) travel again */
    memset(&o, 0, sizeof o);
    rc = vmig_migrate_tree(src, dst, table, NULL, &o, &st);
    if (rc != VMIG_OK || st.blocks_skipped == 0 || st.files_untrusted != 0 || st.bytes_d2h >= st.bytes_total) { printf("diff pass rc=%d skipped=%llu of %llu: %s\n", rc,
        (unsigned long long)st.blocks_skipped, (unsigned long long)st.blocks_total, vmig_last_error()); return 1; }
    printf("diff ok skipped=%llu of %llu\n", (unsigned long long)st.blocks_skipped, (unsigned long long)st.blocks_total);

    /* a bad call sets the message for THIS thread only */
    rc = vmig_migrate_tree(src, "/nonexistent/vmig/dst", NULL, NULL, NULL, NULL);
    if (rc != VMIG_EIO || strlen(vmig_last_error()) == 0) { printf("bad dst rc=%d\n", rc); return 1; }

    /* moveVolumeData: dst -> moved, source entries unlinked */
    rc = vmig_move_dir(dst, moved);
    if (rc != VMIG_OK) { printf("move failed rc=%d: %s\n", rc, vmig_last_error()); return 1; }
    if (vmig_dir_size(dst, &bytes, &nf) != VMIG_OK || nf != 0) { printf("move left %llu files behind\n", (unsigned long long)nf); return 1; }
    if (vmig_dir_size(moved, &bytes, &nf) != VMIG_OK || bytes != src_dir_size) { printf("moved size %lld != %lld\n", (long long)bytes, (long long)src_dir_size); return 1; }
    printf("move ok\n");
    vmig_shutdown();
    return 0;
}
