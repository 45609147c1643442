// Process-wide levins context shared by the host mirror classes (one GPU per process, like the rest of the stack).
#ifndef LEVINS_HOST_DEVICE_H
#define LEVINS_HOST_DEVICE_H

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "levins/abi.h"

namespace levins_host {

inline int &device_ordinal() {
    static int dev = [] {
        const char *e = std::getenv("LEVINS_DEVICE");
        return e ? std::atoi(e) : 0;
    }();
    return dev;
}

// Throws when the library cannot run: there is no CPU fallback behind this surface.
inline levins_ctx *context(const levins_config *cfg = nullptr) {
    static levins_ctx *ctx = nullptr;
    if (!ctx) {
        const int st = levins_create(device_ordinal(), cfg, &ctx);
        if (st != LEVINS_OK)
            throw std::runtime_error(std::string("levins_create failed: ") + levins_status_string(st) +
                                     " (an sm_100 GPU is required; no CPU fallback)");
    }
    return ctx;
}

inline void check(int st, const char *what) {
    if (st < 0) throw std::runtime_error(std::string(what) + ": " + levins_last_error(context()));
}

} // namespace levins_host
#endif
