/* Plain-C consumer of include/b200dct.h: proves the drop-in boundary is a C ABI (compiles as
 * C99 with -pedantic -Werror, links against libb200dct.so, no C++ or torch types anywhere).
 * Runs without a GPU: exercises the host-only entry points and checks that compute entry
 * points refuse to run (there is no CPU fallback).  Exit code 0 = all checks passed. */
#include <stdio.h>
#include <string.h>

#include "b200dct.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "abi_c99: check failed: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(int argc, char **argv)
{
    b200dct_plan *plan = NULL;
    float q[64], q2[64], t[64];
    static float img[64 * 4], out[64 * 4];
    int i, rc, expect_device = argc > 1 && strcmp(argv[1], "--device") == 0;

    CHECK(b200dct_version() == B200DCT_VERSION);
    CHECK(b200dct_plan_create(&plan) == B200DCT_OK && plan != NULL);
    CHECK(b200dct_plan_is_sparse(plan) == 1);
    CHECK(b200dct_plan_get_quant(plan, q) == B200DCT_OK && q[0] == 16.0f && q[63] == 99.0f);
    for (i = 0; i < 64; i++) q2[i] = q[i] * 2.0f;
    CHECK(b200dct_plan_set_quant(plan, q2) == B200DCT_OK);
    q2[5] = 0.0f;
    CHECK(b200dct_plan_set_quant(plan, q2) == B200DCT_ERR_QUANT);
    for (i = 0; i < 64; i++) t[i] = (i % 9 == 0) ? 1.0f : 0.0f; /* identity: a dense T */
    CHECK(b200dct_plan_set_transform(plan, t) == B200DCT_OK && b200dct_plan_is_sparse(plan) == 0);
    CHECK(b200dct_zigzag_mask(1) == 1u && b200dct_zigzag_mask(3) == ((1u << 0) | (1u << 1) | (1u << 8)));
    CHECK(b200dct_plan_set_keep_mask(plan, b200dct_zigzag_mask(10)) == B200DCT_OK);
    CHECK(b200dct_plan_set_path(plan, B200DCT_PATH_DIRECT) == B200DCT_OK);
    CHECK(b200dct_plan_set_path(plan, (b200dct_path)7) == B200DCT_ERR_ARG);
    CHECK(strlen(b200dct_error_string(B200DCT_ERR_SHAPE)) > 0);
    CHECK(b200dct_metrics_workspace_bytes(64, 64) >= 3 * sizeof(double));
    /* argument errors come before any device work */
    CHECK(b200dct_roundtrip(plan, img, B200DCT_F32, 64, out, B200DCT_F32, 64, NULL, B200DCT_F32, 0, 8, 12, NULL) == B200DCT_ERR_SHAPE);
    CHECK(b200dct_roundtrip(NULL, img, B200DCT_F32, 64, out, B200DCT_F32, 64, NULL, B200DCT_F32, 0, 8, 16, NULL) == B200DCT_ERR_ARG);
    /* valid arguments: with a device this would need device pointers, so only the no-device case is run */
    if (!expect_device) {
        rc = b200dct_roundtrip_host(plan, img, B200DCT_F32, out, B200DCT_F32, 16, 16);
        CHECK(rc != B200DCT_OK); /* no CUDA device: refused, never computed on the CPU */
    }
    b200dct_plan_destroy(plan);
    printf("abi_c99 ok\n");
    return 0;
}
