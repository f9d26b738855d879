// compat.cu -- the reference's host entry points on top of the C ABI (see
// include/b200dct_compat.h for the contract and the reference file:line of each item).
#include "b200dct_compat.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// CHECK_CUDA contract of the reference (main_newAppr.cu:9-17): print and exit.
#define COMPAT_CHECK(call)                                                        \
    do {                                                                          \
        int err__ = (int)(call);                                                  \
        if (err__ != 0) {                                                         \
            printf("%s : %d", b200dct_error_string(err__), __LINE__);             \
            exit(EXIT_FAILURE);                                                   \
        }                                                                         \
    } while (0)

namespace {

struct CompatState {
    b200dct_plan *plan = nullptr;
    float last_T[64];
    bool have_T = false;
    bool cache_T = false;          // opt-in: trust the device pointer, skip the per-call fetch
    const float *cached_ptr = nullptr;
    float *d_q = nullptr;          // device copy of Q for cublasDCTv2's in-place dequantisation
    float d_q_host[64];
    bool d_q_valid = false;
    int sms = 0;
    bool side_effects = true;
    bool print_timing = true;
    float last_ms = 0.0f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
thread_local CompatState g;

b200dct_plan *plan()
{
    if (!g.plan) COMPAT_CHECK(b200dct_plan_create(&g.plan));
    return g.plan;
}

// The reference takes T as a device pointer on every call (main_newAppr.cu:99).  Fetch
// its 256 bytes and re-plan only when the contents change.
void sync_transform(const float *d_T)
{
    // b200dct_compat_cache_transform(1): the caller promises that the 64 floats behind a given device
    // pointer do not change between calls (true for every reference program: T is uploaded once,
    // main_newAppr.cu:88-95), so the blocking 256-byte D2H copy in front of every call is skipped
    if (g.cache_T && g.have_T && d_T == g.cached_ptr) return;
    g.cached_ptr = d_T;
    float t[64];
    COMPAT_CHECK(cudaMemcpy(t, d_T, sizeof(t), cudaMemcpyDeviceToHost));
    if (!g.have_T || memcmp(t, g.last_T, sizeof(t)) != 0) {
        COMPAT_CHECK(b200dct_plan_set_transform(plan(), t));
        memcpy(g.last_T, t, sizeof(t));
        g.have_T = true;
    }
}

void timer_start()
{
    if (!g.ev0) {
        COMPAT_CHECK(cudaEventCreate(&g.ev0));
        COMPAT_CHECK(cudaEventCreate(&g.ev1));
    }
    COMPAT_CHECK(cudaEventRecord(g.ev0, 0));
}

void timer_stop(const char *what, int W, int H)
{
    COMPAT_CHECK(cudaEventRecord(g.ev1, 0));
    COMPAT_CHECK(cudaEventSynchronize(g.ev1));
    COMPAT_CHECK(cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1));
    COMPAT_CHECK(cudaGetLastError());
    if (g.print_timing) printf("%s (%d,%d): %f ms\n", what, W, H, g.last_ms); // main_newAppr.cu:287,328
}

void forward(float *image, int H, int W, const float *d_T, float *result)
{
    sync_transform(d_T);
    const size_t pitch = (size_t)W * sizeof(float);
    timer_start();
    COMPAT_CHECK(b200dct_forward(plan(), image, B200DCT_F32, pitch, result, B200DCT_F32, pitch,
                                 g.side_effects ? image : nullptr, H, W, nullptr));
    timer_stop("DCT", W, H);
}

__global__ void k_dequant_inplace(float *c, const float *q, int W, size_t n)
{
    // multiply_matrices in place (main_cublass_2.cu:282): c *= Q[(y%8)*8 + x%8]
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t y = i / W, x = i - y * W;
        c[i] *= q[(y & 7) * 8 + (x & 7)];
    }
}

// dequant_in_place: cublasDCTv2 leaves its coefficient buffer dequantised (main_cublass_2.cu:282).  The
// transform reads the ORIGINAL coefficients in one fused pass; the in-place product is then written by
// a small element-wise kernel in the same stream, inside the timed and synchronised region, so the
// caller observes the same buffer contents as with the reference when the call returns.
void inverse(const float *coef, int H, int W, const float *d_T, float *result, float *dequant_in_place = nullptr)
{
    sync_transform(d_T);
    const size_t pitch = (size_t)W * sizeof(float);
    if (dequant_in_place) {
        float q[64];
        COMPAT_CHECK(b200dct_plan_get_quant(plan(), q));
        if (!g.d_q) COMPAT_CHECK(cudaMalloc(&g.d_q, sizeof(q)));
        if (!g.d_q_valid || memcmp(q, g.d_q_host, sizeof(q)) != 0) { // upload only when the table changed
            COMPAT_CHECK(cudaMemcpy(g.d_q, q, sizeof(q), cudaMemcpyHostToDevice));
            memcpy(g.d_q_host, q, sizeof(q));
            g.d_q_valid = true;
        }
        if (!g.sms) {
            int dev = 0;
            COMPAT_CHECK(cudaGetDevice(&dev));
            COMPAT_CHECK(cudaDeviceGetAttribute(&g.sms, cudaDevAttrMultiProcessorCount, dev));
        }
    }
    timer_start();
    COMPAT_CHECK(b200dct_inverse(plan(), coef, B200DCT_F32, pitch, result, B200DCT_F32, pitch, H, W, nullptr));
    if (dequant_in_place) {
        const size_t n = (size_t)H * W;
        size_t blocks = (n + 255) / 256;
        if (blocks > (size_t)g.sms * 8) blocks = (size_t)g.sms * 8;
        k_dequant_inplace<<<(unsigned)blocks, 256>>>(dequant_in_place, g.d_q, W, n);
    }
    timer_stop("IDCT", W, H);
}

} // namespace

void dct_all_blocks_cuda(float *image_matrix, const int img_height, const int img_width,
                         const float *transform_matrix, float *result)
{
    forward(image_matrix, img_height, img_width, transform_matrix, result);
}

void idct_all_blocks_cuda(const float *image_matrix, const int img_height, const int img_width,
                          const float *transform_matrix, float *result)
{
    inverse(image_matrix, img_height, img_width, transform_matrix, result);
}

void dct_all_blocks(float *image_matrix, int img_height, int img_width, const float *transform_matrix,
                    float *result, cublasHandle_t)
{
    forward(image_matrix, img_height, img_width, transform_matrix, result);
}

void idct_all_blocks(const float *image_matrix, int img_height, int img_width, const float *transform_matrix,
                     float *result, cublasHandle_t)
{
    inverse(image_matrix, img_height, img_width, transform_matrix, result);
}

// cublasDCTv2 (non-const input): see inverse()
void idct_all_blocks(float *image_matrix, int img_height, int img_width, const float *transform_matrix,
                     float *result, cublasHandle_t)
{
    inverse(image_matrix, img_height, img_width, transform_matrix, result, g.side_effects ? image_matrix : nullptr);
}

extern "C" {

int b200dct_compat_set_quant(const float *q)
{
    return b200dct_plan_set_quant(plan(), q);
}
int b200dct_compat_set_keep_mask(uint64_t mask) { return b200dct_plan_set_keep_mask(plan(), mask); }
void b200dct_compat_set_options(int side_effects, int print_timing)
{
    g.side_effects = side_effects != 0;
    g.print_timing = print_timing != 0;
}
void b200dct_compat_cache_transform(int on)
{
    g.cache_T = on != 0;
    g.cached_ptr = nullptr; // also the way to invalidate: the next call fetches T again
}
float b200dct_compat_last_ms(void) { return g.last_ms; }
void b200dct_compat_dct(float *image, int H, int W, const float *d_T, float *result) { dct_all_blocks_cuda(image, H, W, d_T, result); }
void b200dct_compat_idct(const float *coef, int H, int W, const float *d_T, float *result) { idct_all_blocks_cuda(coef, H, W, d_T, result); }
void b200dct_compat_idct_inplace_dequant(float *coef, int H, int W, const float *d_T, float *result)
{
    idct_all_blocks(coef, H, W, d_T, result, (cublasHandle_t) nullptr);
}
}
