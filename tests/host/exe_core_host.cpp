// Host harness for kanzi-go_b200/csrc/kz_exe_core.cuh (tests/test_exe_core_host.py). A test of host/device-shared logic.
#include <cstdint>

#include "../../kanzi-go_b200/csrc/kz_exe_core.cuh"

extern "C" {
int64_t xc_forward(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap, int dt_in, int* dt_out) {
    int dt = dt_in;
    const int64_t r = kz::exec::forward(src, n, dst, cap, &dt);
    *dt_out = dt;
    return r;
}
int64_t xc_inverse(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap) { return kz::exec::inverse(src, n, dst, cap); }
}
