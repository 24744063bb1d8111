/* Plain-C client of include/vcs_b200.h: proves the boundary is a C ABI (no Python, no torch).
 * Built and run by tests/test_gpu_abi_c.py on the GPU box:  abi_smoke <path to libvcs_b200.so>
 * Searches a synthetic pair with the reference's own parameters and with a +/-8 full search,
 * compresses and decompresses the residual, prints a few checksums the Python test compares
 * with the oracle. */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/vcs_b200.h"

#define SYM(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    void *lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    SYM(vcs_create) SYM(vcs_destroy) SYM(vcs_last_error) SYM(vcs_me_reference_params) SYM(vcs_me_fullsearch_params)
    SYM(vcs_num_blocks) SYM(vcs_me_search_host) SYM(vcs_mc_host) SYM(vcs_sub_wrap_host) SYM(vcs_compress_host)
    SYM(vcs_decompress_host)
    enum { H = 64, W = 96 };
    static uint8_t ref[H * W * 3], cur[H * W * 3], pred[H * W * 3], resid[H * W * 3], out[H * W * 3];
    uint32_t s = 2463534242u;   /* xorshift: the Python side regenerates the same frames */
    for (int i = 0; i < H * W * 3; ++i) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; ref[i] = (uint8_t)(s >> 11); }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < 3; ++c) {
                int sy = y + 2 < H ? y + 2 : y, sx = x >= 3 ? x - 3 : x;
                cur[(y * W + x) * 3 + c] = ref[(sy * W + sx) * 3 + c];
            }
    vcs_ctx *ctx = NULL;
    if (p_vcs_create(0, &ctx) != VCS_OK) { fprintf(stderr, "vcs_create failed (no GPU?)\n"); return 3; }
    for (int pass = 0; pass < 2; ++pass) {
        vcs_me_params p;
        if (pass == 0) p_vcs_me_reference_params(H, W, 8, &p);
        else p_vcs_me_fullsearch_params(H, W, 16, 8, VCS_METRIC_SAD, -1, &p);
        const int N = p_vcs_num_blocks(H, W, p.bs);
        int16_t *mv = (int16_t *)malloc(sizeof(int16_t) * 2 * N);
        uint32_t *cost = (uint32_t *)malloc(sizeof(uint32_t) * N);
        uint8_t *flags = (uint8_t *)malloc(N);
        if (p_vcs_me_search_host(ctx, &p, cur, ref, mv, cost, flags) != VCS_OK) { fprintf(stderr, "%s\n", p_vcs_last_error(ctx)); return 4; }
        long long smv = 0, sc = 0; int sf = 0;
        for (int k = 0; k < N; ++k) { smv += (long long)mv[2 * k] * 131 + mv[2 * k + 1] * 7 + k * (mv[2 * k] ^ mv[2 * k + 1]); sc += cost[k] % 1000003u; sf += flags[k]; }
        printf("pass %d N %d mvsum %lld costsum %lld flagsum %d\n", pass, N, smv, sc, sf);
        if (pass == 1) {
            if (p_vcs_mc_host(ctx, H, W, p.bs, ref, mv, pred) != VCS_OK) return 5;
            if (p_vcs_sub_wrap_host(ctx, cur, pred, (size_t)H * W * 3, resid) != VCS_OK) return 5;
            double *planes = (double *)malloc(sizeof(double) * 3 * H * W);
            if (p_vcs_compress_host(ctx, H, W, resid, VCS_COEF_F64, planes) != VCS_OK) return 6;
            if (p_vcs_decompress_host(ctx, H, W, VCS_COEF_F64, planes, pred, out) != VCS_OK) return 6;
            double ps = 0; long long os = 0;
            for (int k = 0; k < 3 * H * W; ++k) ps += planes[k] * ((k % 17) + 1);
            for (int k = 0; k < H * W * 3; ++k) os += out[k] * ((k % 13) + 1);
            printf("planesum %.17g outsum %lld\n", ps, os);
            free(planes);
        }
        free(mv); free(cost); free(flags);
    }
    p_vcs_destroy(ctx);
    return 0;
}
