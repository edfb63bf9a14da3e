// vmig_common.h -- shared host-side helpers of libvmig (errors, timing, small utilities).
#pragma once
#include <stdint.h>
#include <string>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cerrno>
#include <cstring>
#include "../../include/vmig.h"

namespace vmig {

// Thread-local detail of the last failure on this thread (vmig_last_error()).
void set_last_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void set_last_error_str(const std::string& s);
const char* last_error_cstr();

inline uint64_t now_ns() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    set_last_error_str(buf);
    return code;
}

inline std::string errno_str(int e) { char b[128]; return std::string(strerror_r(e, b, sizeof b)); }

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

long env_long(const char* name, long dflt);

}  // namespace vmig
