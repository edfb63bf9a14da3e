// vmig_cufile.cpp -- dlopen() wrapper around libcufile (see vmig_cufile.h).  New code: the reference has no storage
// path beyond `tar` reading through the page cache (utils/copy.go:17-27).
#include "vmig_cufile.h"
#include "vmig_common.h"

#include <cufile.h>
#include <dlfcn.h>
#include <mutex>

namespace vmig {

namespace {
struct Api {
    void* lib = nullptr;
    CUfileError_t (*DriverOpen)() = nullptr;
    CUfileError_t (*DriverClose)() = nullptr;
    CUfileError_t (*HandleRegister)(CUfileHandle_t*, CUfileDescr_t*) = nullptr;
    void (*HandleDeregister)(CUfileHandle_t) = nullptr;
    CUfileError_t (*BufRegister)(const void*, size_t, int) = nullptr;
    CUfileError_t (*BufDeregister)(const void*) = nullptr;
    ssize_t (*Read)(CUfileHandle_t, void*, size_t, off_t, off_t) = nullptr;
    ssize_t (*Write)(CUfileHandle_t, const void*, size_t, off_t, off_t) = nullptr;
    bool driver_open = false;
};
Api g_api;
std::mutex g_mu;
int g_state = 0;          // 0 untried, 1 ok, -1 failed
std::string g_why;

template <class F> bool sym(void* lib, const char* name, F* out) { *out = reinterpret_cast<F>(dlsym(lib, name)); return *out != nullptr; }
}  // namespace

int cufile_open()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_state == 1) return VMIG_OK;
    if (g_state == -1) return fail(VMIG_EINVAL, "GPUDirect Storage unavailable: %s", g_why.c_str());
    const char* names[] = {"libcufile.so.0", "libcufile.so", "/usr/local/cuda/lib64/libcufile.so.0", "/usr/local/cuda/lib64/libcufile.so"};
    for (const char* n : names) { g_api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (g_api.lib) break; }
    if (!g_api.lib) { g_state = -1; g_why = std::string("dlopen libcufile.so: ") + (dlerror() ? dlerror() : "not found"); return fail(VMIG_EINVAL, "GPUDirect Storage unavailable: %s", g_why.c_str()); }
    bool ok = sym(g_api.lib, "cuFileDriverOpen", &g_api.DriverOpen) && sym(g_api.lib, "cuFileHandleRegister", &g_api.HandleRegister) &&
              sym(g_api.lib, "cuFileHandleDeregister", &g_api.HandleDeregister) && sym(g_api.lib, "cuFileBufRegister", &g_api.BufRegister) &&
              sym(g_api.lib, "cuFileBufDeregister", &g_api.BufDeregister) && sym(g_api.lib, "cuFileRead", &g_api.Read) &&
              sym(g_api.lib, "cuFileWrite", &g_api.Write);
    if (!sym(g_api.lib, "cuFileDriverClose_v2", &g_api.DriverClose)) sym(g_api.lib, "cuFileDriverClose", &g_api.DriverClose);
    if (!ok) { g_state = -1; g_why = "libcufile.so lacks an expected symbol"; return fail(VMIG_EINVAL, "GPUDirect Storage unavailable: %s", g_why.c_str()); }
    const CUfileError_t e = g_api.DriverOpen();
    if (e.err != CU_FILE_SUCCESS) {
        g_state = -1; g_why = std::string("cuFileDriverOpen: ") + CUFILE_ERRSTR(e.err);
        return fail(VMIG_EINVAL, "GPUDirect Storage unavailable: %s", g_why.c_str());
    }
    g_api.driver_open = true; g_state = 1;
    return VMIG_OK;
}

bool cufile_loaded() { std::lock_guard<std::mutex> lk(g_mu); return g_state == 1; }

int cufile_handle_open(int fd, void** handle)
{
    CUfileDescr_t d; memset(&d, 0, sizeof d);
    d.type = CU_FILE_HANDLE_TYPE_OPAQUE_FD; d.handle.fd = fd;
    CUfileHandle_t h = nullptr;
    const CUfileError_t e = g_api.HandleRegister(&h, &d);
    if (e.err != CU_FILE_SUCCESS) return fail(VMIG_EIO, "cuFileHandleRegister: %s", CUFILE_ERRSTR(e.err));
    *handle = h;
    return VMIG_OK;
}
void cufile_handle_close(void* handle) { if (handle && g_api.HandleDeregister) g_api.HandleDeregister((CUfileHandle_t)handle); }

void cufile_buf_register(void* dev_ptr, size_t bytes) { if (g_api.BufRegister) (void)g_api.BufRegister(dev_ptr, bytes, 0); }
void cufile_buf_deregister(void* dev_ptr) { if (g_api.BufDeregister) (void)g_api.BufDeregister(dev_ptr); }

ssize_t cufile_read(void* handle, void* dev_base, size_t bytes, off_t off, off_t dev_off)
{
    const ssize_t r = g_api.Read((CUfileHandle_t)handle, dev_base, bytes, off, dev_off);
    if (r >= 0) return r;
    if (r == -1) return fail(VMIG_EIO, "cuFileRead: %s", errno_str(errno).c_str());
    return fail(VMIG_EIO, "cuFileRead: %s", CUFILE_ERRSTR((int)-r));
}
ssize_t cufile_write(void* handle, const void* dev_base, size_t bytes, off_t off, off_t dev_off)
{
    const ssize_t r = g_api.Write((CUfileHandle_t)handle, dev_base, bytes, off, dev_off);
    if (r >= 0) return r;
    if (r == -1) return fail(VMIG_EIO, "cuFileWrite: %s", errno_str(errno).c_str());
    return fail(VMIG_EIO, "cuFileWrite: %s", CUFILE_ERRSTR((int)-r));
}

void cufile_shutdown()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_state == 1 && g_api.driver_open && g_api.DriverClose) { g_api.DriverClose(); g_api.driver_open = false; }
    g_state = 0;            // the library handle stays loaded; DriverOpen runs again on the next use
}

}  // namespace vmig
