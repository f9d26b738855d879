/*
 * dct_oracle.c -- CPU restatement of the reference's 8x8 block-transform path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package, the C-ABI
 * library, the compat wrappers) may import, link or execute this file.  It is used
 * by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / reference
 * legs as the *checker* and as the sequential CPU baseline.
 *
 * Parity status: the reference ships NO tests, golden vectors or fixtures
 * (SURVEY.md section 4).  This restatement is pinned instead by running the
 * reference's own kernels, compiled unmodified from /root/reference into
 * oracle/_ref/ (see oracle/Makefile), on a B200 and comparing bit-for-bit
 * (tests/test_gpu_reference.py), and by the committed fixtures in tests/golden/.
 *
 * Each function cites the reference file:line it follows.  Arithmetic contract
 * (SURVEY.md Appendix A): all values are IEEE binary32; every inner product is an
 * ordered chain of 8 fused multiply-adds starting from +0.0f with the summation
 * index ascending (that is what nvcc emits for the reference's `sums += a*b` loops
 * under its default -fmad=true); quantisation is a correctly-rounded division
 * followed by round-half-away-from-zero; dequantisation is a plain multiply.
 *
 * Build: gcc -O2 -ffp-contract=off (so that ONLY the explicit fmaf() calls fuse).
 * The hot loops are cloned for FMA-capable CPUs via target_clones, the default
 * clone falls back to glibc's correctly-rounded software fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BS 8

#if defined(__x86_64__) && defined(__GNUC__)
#define ORACLE_CLONES __attribute__((target_clones("fma", "default")))
#else
#define ORACLE_CLONES
#endif

/* Haweel's approximate-DCT matrix exactly as the reference spells it
 * (main_newAppr.cu:73-81): double literals narrowed to float. */
static const float k_haweel_T[64] = {
    0.35355339, 0.35355339, 0.35355339, 0.35355339, 0.35355339, 0.35355339, 0.35355339, 0.35355339,
    0.5, 0.5, 0, 0, 0, 0, -0.5, -0.5,
    0.4472136, 0.2236068, -0.2236068, -0.4472136, -0.4472136, -0.2236068, 0.2236068, 0.4472136,
    0, 0, -0.70710678, 0, 0, 0.70710678, 0, 0,
    0.35355339, -0.35355339, -0.35355339, 0.35355339, 0.35355339, -0.35355339, -0.35355339, 0.35355339,
    0.5, -0.5, 0, 0, 0, 0, 0.5, -0.5,
    0.2236068, -0.4472136, 0.4472136, -0.2236068, -0.2236068, 0.4472136, -0.4472136, 0.2236068,
    0, 0, 0, -0.70710678, 0.70710678, 0, 0, 0};

/* JPEG luminance quantisation table (main_newAppr.cu:60-68). */
static const float k_jpeg_Q[64] = {
    16, 11, 10, 16, 24, 40, 51, 61,
    12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77,
    24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101,
    72, 92, 95, 98, 112, 100, 103, 99};

/* JPEG zig-zag scan order as (row*8+col); the retained-coefficient mask keeps the
 * first k entries (README.md:63 of the reference). */
static const unsigned char k_zigzag[64] = {
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5,
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

const float *oracle_haweel_T(void) { return k_haweel_T; }
const float *oracle_jpeg_Q(void) { return k_jpeg_Q; }

/* True orthonormal DCT-II matrix in float ("exact DCT" = the dense variants fed a
 * real DCT matrix).  Computed in double, narrowed once. */
void oracle_dct2_T(float *T)
{
    for (int k = 0; k < BS; k++)
        for (int n = 0; n < BS; n++) {
            double c = (k == 0) ? sqrt(1.0 / BS) : sqrt(2.0 / BS);
            T[k * BS + n] = (float)(c * cos((2 * n + 1) * k * M_PI / (2.0 * BS)));
        }
}

/* Bit i of the mask <=> coefficient at (row*8+col)==i is kept. */
uint64_t oracle_zigzag_mask(int k)
{
    uint64_t m = 0;
    if (k >= 64) return ~(uint64_t)0;
    for (int i = 0; i < k; i++) m |= (uint64_t)1 << k_zigzag[i];
    return m;
}

/* Compact coefficient stream (not in the reference, whose coefficient
 * plane is a float image, main_newAppr.cu:99-103): block-major, block (r,c) -> 64 consecutive
 * int16 in zig-zag order, values = the integer-valued float coefficients saturated to int16. */
void oracle_zigzag_i16(const float *coef, int H, int W, int16_t *out)
{
    const int bx = W / BS;
    for (int r = 0; r < H / BS; r++)
        for (int c = 0; c < bx; c++)
            for (int k = 0; k < 64; k++) {
                float v = coef[(size_t)(r * BS + k_zigzag[k] / BS) * W + c * BS + k_zigzag[k] % BS];
                if (v > 32767.0f) v = 32767.0f;
                if (v < -32768.0f) v = -32768.0f;
                out[((size_t)r * bx + c) * 64 + k] = (int16_t)lrintf(v);
            }
}
/* and back: the float coefficient plane the inverse transform starts from */
void oracle_unzigzag_i16(const int16_t *in, int H, int W, float *coef)
{
    const int bx = W / BS;
    for (int r = 0; r < H / BS; r++)
        for (int c = 0; c < bx; c++)
            for (int k = 0; k < 64; k++)
                coef[(size_t)(r * BS + k_zigzag[k] / BS) * W + c * BS + k_zigzag[k] % BS] =
                    (float)in[((size_t)r * bx + c) * 64 + k];
}

/* Reference input generator: srand(seed); img[i*N+j] = rand()%256
 * (Benchmark_code/benchmark_fastAppr.cu:44-47).  glibc rand(). */
void oracle_fill_rand(float *img, size_t n, unsigned seed)
{
    srand(seed);
    for (size_t i = 0; i < n; i++) img[i] = (float)(rand() % 256);
}

void oracle_fill_rand_u8(unsigned char *img, size_t n, unsigned seed)
{
    srand(seed);
    for (size_t i = 0; i < n; i++) img[i] = (unsigned char)(rand() % 256);
}

/* utils.cu:10-15 */
void oracle_convert_to_float(const unsigned char *in, float *out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = (float)in[i];
}

/* utils.cu:18-24 : clamp to [0,255] then C cast (truncation toward zero). */
void oracle_convert_to_u8(const float *in, unsigned char *out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = (unsigned char)fminf(fmaxf(in[i], 0.0f), 255.0f);
}

/* ---- one block-row strip of the forward path -------------------------------------
 * sub_matrix_scalar   utils_kernels.cu:8-18     X = img - 128
 * cuda_matrix_dct     main_newAppr.cu:177-211   M = T.X (:193-197), Y = M.T^T (:206-209)
 * divide_matrices     utils_kernels.cu:34-44    C = round(Y / Q[ty*8+tx])
 * `keep` is applied to the quantised coefficients (exact +0.0f where dropped).
 */
ORACLE_CLONES
static void fwd_block(const float *src, size_t pitch, const float *T, const float *Q,
                      uint64_t keep, float *dst, size_t dpitch, float *shifted, size_t spitch)
{
    float X[BS][BS], M[BS][BS];
    for (int y = 0; y < BS; y++)
        for (int x = 0; x < BS; x++) {
            X[y][x] = src[y * pitch + x] - 128.0f;
            if (shifted) shifted[y * spitch + x] = X[y][x];
        }
    for (int y = 0; y < BS; y++)
        for (int x = 0; x < BS; x++) {
            float s = 0.0f;
            for (int i = 0; i < BS; i++) s = fmaf(T[y * BS + i], X[i][x], s);
            M[y][x] = s;
        }
    for (int y = 0; y < BS; y++)
        for (int x = 0; x < BS; x++) {
            float s = 0.0f;
            for (int i = 0; i < BS; i++) s = fmaf(M[y][i], T[x * BS + i], s);
            float c = roundf(s / Q[y * BS + x]);
            if (!((keep >> (y * BS + x)) & 1)) c = 0.0f;
            dst[y * dpitch + x] = c;
        }
}

/* ---- inverse path -----------------------------------------------------------------
 * multiply_matrices   utils_kernels.cu:47-57    D = C * Q[ty*8+tx]
 * cuda_matrix_idct    main_newAppr.cu:220-250   M = T^T.D (:236-239), R = M.T (:246-248)
 * add_matrix_scalar   utils_kernels.cu:21-31    out = R + 128   (not clamped)
 */
ORACLE_CLONES
static void inv_block(const float *src, size_t pitch, const float *T, const float *Q,
                      float *dst, size_t dpitch)
{
    float D[BS][BS], M[BS][BS];
    for (int y = 0; y < BS; y++)
        for (int x = 0; x < BS; x++) D[y][x] = src[y * pitch + x] * Q[y * BS + x];
    for (int y = 0; y < BS; y++)
        for (int x = 0; x < BS; x++) {
            float s = 0.0f;
            for (int i = 0; i < BS; i++) s = fmaf(T[i * BS + y], D[i][x], s);
            M[y][x] = s;
        }
    for (int y = 0; y < BS; y++)
        for (int x = 0; x < BS; x++) {
            float s = 0.0f;
            for (int i = 0; i < BS; i++) s = fmaf(M[y][i], T[i * BS + x], s);
            dst[y * dpitch + x] = s + 128.0f;
        }
}

/* dct_all_blocks_cuda (main_newAppr.cu:252-291).  If `shifted` is non-NULL it
 * receives img-128, the value the reference leaves in its (mutated) input buffer
 * (:273).  threads<=1 -> sequential. */
void oracle_dct(const float *img, int H, int W, const float *T, const float *Q,
                uint64_t keep, float *coef, float *shifted, int threads)
{
    const int by = H / BS, bx = W / BS;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
    for (int r = 0; r < by; r++)
        for (int c = 0; c < bx; c++) {
            size_t off = (size_t)r * BS * W + (size_t)c * BS;
            fwd_block(img + off, W, T, Q, keep, coef + off, W,
                      shifted ? shifted + off : NULL, W);
        }
    (void)threads;
}

/* idct_all_blocks_cuda (main_newAppr.cu:293-332). */
void oracle_idct(const float *coef, int H, int W, const float *T, const float *Q,
                 float *out, int threads)
{
    const int by = H / BS, bx = W / BS;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
    for (int r = 0; r < by; r++)
        for (int c = 0; c < bx; c++) {
            size_t off = (size_t)r * BS * W + (size_t)c * BS;
            inv_block(coef + off, W, T, Q, out + off, W);
        }
    (void)threads;
}

/* Fused round trip on float pixels: coef (optional) and reconstructed pixels. */
void oracle_roundtrip(const float *img, int H, int W, const float *T, const float *Q,
                      uint64_t keep, float *coef_or_null, float *out, int threads)
{
    const int by = H / BS, bx = W / BS;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
    for (int r = 0; r < by; r++)
        for (int c = 0; c < bx; c++) {
            float tmp[BS * BS];
            size_t off = (size_t)r * BS * W + (size_t)c * BS;
            fwd_block(img + off, W, T, Q, keep, tmp, BS, NULL, 0);
            if (coef_or_null)
                for (int y = 0; y < BS; y++)
                    memcpy(coef_or_null + off + (size_t)y * W, tmp + y * BS, BS * sizeof(float));
            inv_block(tmp, BS, T, Q, out + off, W);
        }
    (void)threads;
}

/* u8 in -> u8 out round trip: convertToFloat, forward, inverse, convertToUnsignedChar
 * (main_newAppr.cu:47,99,120,141). */
void oracle_roundtrip_u8(const unsigned char *img, int H, int W, const float *T,
                         const float *Q, uint64_t keep, float *coef_or_null,
                         unsigned char *out, int threads)
{
    const int by = H / BS, bx = W / BS;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 1 ? threads : 1)
#endif
    for (int r = 0; r < by; r++)
        for (int c = 0; c < bx; c++) {
            float blk[BS * BS], tmp[BS * BS], rec[BS * BS];
            size_t off = (size_t)r * BS * W + (size_t)c * BS;
            for (int y = 0; y < BS; y++)
                oracle_convert_to_float(img + off + (size_t)y * W, blk + y * BS, BS);
            fwd_block(blk, BS, T, Q, keep, tmp, BS, NULL, 0);
            if (coef_or_null)
                for (int y = 0; y < BS; y++)
                    memcpy(coef_or_null + off + (size_t)y * W, tmp + y * BS, BS * sizeof(float));
            inv_block(tmp, BS, T, Q, rec, BS);
            for (int y = 0; y < BS; y++)
                oracle_convert_to_u8(rec + y * BS, out + off + (size_t)y * W, BS);
        }
    (void)threads;
}

/* MSE and PEEN as recovered from the reference's README table:
 * MSE = sum((x-y)^2)/N, PEEN% = 100*sqrt(sum((x-y)^2)/sum(x^2)); double accumulation. */
void oracle_metrics_u8(const unsigned char *x, const unsigned char *y, size_t n,
                       double *mse, double *peen)
{
    double se = 0.0, e = 0.0;
    for (size_t i = 0; i < n; i++) {
        double d = (double)x[i] - (double)y[i];
        se += d * d;
        e += (double)x[i] * (double)x[i];
    }
    *mse = se / (double)n;
    *peen = e > 0.0 ? 100.0 * sqrt(se / e) : 0.0;
}

void oracle_metrics_f32(const float *x, const float *y, size_t n, double *mse, double *peen)
{
    double se = 0.0, e = 0.0;
    for (size_t i = 0; i < n; i++) {
        double d = (double)x[i] - (double)y[i];
        se += d * d;
        e += (double)x[i] * (double)x[i];
    }
    *mse = se / (double)n;
    *peen = e > 0.0 ? 100.0 * sqrt(se / e) : 0.0;
}

/* FNV-1a-64 over little-endian int32 coefficients / u8 pixels. */
uint64_t oracle_fnv_coef(const float *coef, size_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; i++) {
        int32_t v = (int32_t)coef[i];
        for (int b = 0; b < 4; b++) {
            h ^= (uint64_t)((uint32_t)v >> (8 * b)) & 0xff;
            h *= 0x100000001b3ull;
        }
    }
    return h;
}

uint64_t oracle_fnv_u8(const unsigned char *p, size_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 0x100000001b3ull;
    }
    return h;
}

/* Wall-clock seconds of `reps` round trips (best of reps), for the CPU baseline. */
double oracle_time_roundtrip(const float *img, int H, int W, float *out, int reps, int threads)
{
    double best = 1e30;
    for (int i = 0; i < reps; i++) {
        struct timespec a, b;
        clock_gettime(CLOCK_MONOTONIC, &a);
        oracle_roundtrip(img, H, W, k_haweel_T, k_jpeg_Q, ~(uint64_t)0, NULL, out, threads);
        clock_gettime(CLOCK_MONOTONIC, &b);
        double t = (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
        if (t < best) best = t;
    }
    return best;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
