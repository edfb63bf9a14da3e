// tree_unit.cpp -- CPU unit test of the source-side path safety in csrc/vmig_tree.cpp (compiled and run by
// tests/test_host.py).  The source layer may belong to a running tenant: open_beneath() must refuse symlinks and
// escapes, the walk must not descend through a symlinked directory, remove_source() must not leave the tree.
//   g++ -std=c++17 -I gpu-docker-api_b200/csrc tests/tree_unit.cpp gpu-docker-api_b200/csrc/vmig_tree.cpp -o tree_unit
#include "vmig_tree.h"
#include "vmig_common.h"
#include <fcntl.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace vmig {   // the three error-channel functions normally provided by vmig_engine.cu
static thread_local std::string g_err;
void set_last_error(const char* fmt, ...) { char b[1024]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap); g_err = b; }
void set_last_error_str(const std::string& s) { g_err = s; }
const char* last_error_cstr() { return g_err.c_str(); }
long env_long(const char*, long d) { return d; }
}
using namespace vmig;

#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s   [last error: %s]\n", __LINE__, #c, last_error_cstr()); return 1; } } while (0)
static void put(const std::string& p, const char* s) { FILE* f = fopen(p.c_str(), "w"); fputs(s, f); fclose(f); }
static std::string get(int fd) { char b[64] = {0}; ssize_t n = read(fd, b, 63); close(fd); return std::string(b, n > 0 ? (size_t)n : 0); }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string base = argv[1], root = base + "/root", outside = base + "/outside";
    CHECK(system(("mkdir -p " + root + "/a/b " + root + "/victim " + outside).c_str()) == 0);
    put(root + "/a/b/file", "inside"); put(root + "/victim/f", "inside2"); put(outside + "/f", "SECRET"); put(outside + "/secret", "SECRET");
    CHECK(symlink(outside.c_str(), (root + "/a/link").c_str()) == 0);
    CHECK(symlink("b/file", (root + "/a/sl").c_str()) == 0);
    const int rfd = open(root.c_str(), O_RDONLY | O_DIRECTORY);
    CHECK(rfd >= 0);
    int fd = -1;
    CHECK(open_beneath(rfd, "a/b/file", O_RDONLY, &fd) == VMIG_OK && get(fd) == "inside");
    CHECK(open_beneath(rfd, "a/link/secret", O_RDONLY, &fd) == VMIG_ESRCCHANGED);        // symlinked directory on the way
    CHECK(open_beneath(rfd, "a/sl", O_RDONLY, &fd) == VMIG_ESRCCHANGED);                 // symlink as the last component
    CHECK(open_beneath(rfd, "../outside/secret", O_RDONLY, &fd) != VMIG_OK);             // lexical escape
    CHECK(open_beneath(rfd, "/etc/passwd", O_RDONLY, &fd) == VMIG_EINVAL);               // absolute
    CHECK(open_beneath(rfd, "a/missing", O_RDONLY, &fd) == VMIG_ESRCCHANGED);            // vanished since the walk

    // the same through the component-wise fallback used on kernels without openat2
    CHECK(open_beneath_walk(rfd, "a/b/file", O_RDONLY, &fd) == VMIG_OK && get(fd) == "inside");
    CHECK(open_beneath_walk(rfd, "a/link/secret", O_RDONLY, &fd) == VMIG_ESRCCHANGED);
    CHECK(open_beneath_walk(rfd, "a/sl", O_RDONLY, &fd) == VMIG_ESRCCHANGED);
    CHECK(open_beneath_walk(rfd, "../outside/secret", O_RDONLY, &fd) == VMIG_EINVAL);
    CHECK(open_beneath_walk(rfd, "a/missing", O_RDONLY, &fd) == VMIG_ESRCCHANGED);

    Manifest m;
    CHECK(walk_tree(root, 4096, false, &m) == VMIG_OK);
    CHECK(m.files.size() == 2 && m.symlinks.size() == 2 && m.dirs.size() == 4);          // ".", a, a/b, victim; nothing from outside/
    for (auto& e : m.files) CHECK(e.rel == "a/b/file" || e.rel == "victim/f");

    // the tenant swaps victim/ for a symlink to outside/ after the walk: the file must not be reachable any more,
    // and a move's clean-up must not delete outside/f
    CHECK(rename((root + "/victim").c_str(), (root + "/victim.bak").c_str()) == 0);
    CHECK(symlink(outside.c_str(), (root + "/victim").c_str()) == 0);
    CHECK(open_beneath(rfd, "victim/f", O_RDONLY, &fd) == VMIG_ESRCCHANGED);
    CHECK(remove_source(root, m) != VMIG_OK);
    CHECK(access((outside + "/f").c_str(), F_OK) == 0);
    close(rfd);
    printf("tree unit ok\n");
    return 0;
}
