// poco_b200 -- host-side internals shared by the .cu translation units (not part of the C ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/poco_b200.h"

namespace poco {

extern std::atomic<int64_t> g_launches;
void set_error(const std::string& msg);

#define POCO_CHECK(cond, msg)                                                      \
    do {                                                                           \
        if (!(cond)) {                                                             \
            poco::set_error(std::string(__func__) + ": " + (msg));                 \
            return 1;                                                              \
        }                                                                          \
    } while (0)

#define POCO_CUDA(expr)                                                            \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            poco::set_error(std::string(__func__) + ": " #expr " -> " +            \
                            cudaGetErrorString(_e));                               \
            return 2;                                                              \
        }                                                                          \
    } while (0)

#define POCO_LAUNCHED()                                                            \
    do {                                                                           \
        poco::g_launches.fetch_add(1, std::memory_order_relaxed);                  \
        POCO_CUDA(cudaPeekAtLastError());                                          \
    } while (0)

inline int64_t act_pixels_per_crop(const poco_act& a) { return int64_t(a.H + 2) * (a.W + 2); }
int check_act(const poco_act& a, const char* what);

// per-TU launchers
int conv_tc_launch(const poco_conv* d, cudaStream_t s);
int conv_tc_launch_chain(const poco_conv* segs, int n_segs, int32_t* flags, cudaStream_t s);
int conv_ref_launch(const poco_conv* d, cudaStream_t s);
int64_t conv_flops(const poco_conv* d);
int basic_block64_supported(int H, int W);                                  // bblock64_tc.cu: 64 channels, conv2's weights streamed
int basic_block64_launch(const poco_basic_block* d, cudaStream_t s);

}  // namespace poco
