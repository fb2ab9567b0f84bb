// common.h -- error plumbing shared by the host code and the CUDA launchers.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/voxtral.h"

namespace vox {

// Exception carrying a VOX_E* status; converted to a status + thread-local message at the C ABI.
struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &m) : std::runtime_error(m), code(c) {}
};

inline std::string fmt(const char *f, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    return std::string(buf);
}

[[noreturn]] inline void fail(int32_t code, const std::string &m) { throw Error(code, m); }

void set_last_error(const std::string &m);

#define VOX_CHECK(cond, code, ...)                      \
    do {                                                \
        if (!(cond)) ::vox::fail((code), ::vox::fmt(__VA_ARGS__)); \
    } while (0)

}  // namespace vox
