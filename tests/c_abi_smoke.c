/* tests/c_abi_smoke.c -- the C ABI used from plain C (what cgo does): compiled and run by
 * tests/test_host.py::test_c_abi_from_plain_c.  No GPU needed: host-side entry points only, plus
 * the requirement that a data-path call fails with VMIG_ENOGPU (or succeeds on a GPU box). */
#include <stdio.h>
#include <string.h>
#include "vmig.h"

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    const char *src = argv[1], *dst = argv[2];
    int64_t v = 0;
    if (vmig_to_bytes("20GB", &v) != VMIG_OK || v != 21474836480LL) { printf("to_bytes\n"); return 1; }
    if (vmig_to_bytes("1XB", &v) != VMIG_EINVAL) { printf("to_bytes bad unit\n"); return 1; }
    vmig_stats st;
    if (vmig_manifest(src, 0, 0, NULL, &st) != VMIG_OK) { printf("manifest: %s\n", vmig_last_error()); return 1; }
    int64_t bytes = 0; uint64_t nf = 0;
    if (vmig_dir_size(src, &bytes, &nf) != VMIG_OK) { printf("dir_size: %s\n", vmig_last_error()); return 1; }
    printf("version=%s files=%llu bytes=%llu dir_size=%lld\n", vmig_version(), (unsigned long long)st.files,
           (unsigned long long)st.bytes_total, (long long)bytes);
    int rc = vmig_copy_dir(src, dst);               /* utils.CopyDir(src, dest) */
    if (rc == VMIG_OK) { printf("copy ok\n"); return 0; }
    if (rc == VMIG_ENOGPU) { printf("copy refused: %s | %s\n", vmig_strerror(rc), vmig_last_error()); return 0; }
    printf("copy failed rc=%d: %s\n", rc, vmig_last_error());
    return 1;
}
