// Development harness: times smz::gemm_bf16 with the epilogue variants / shapes the VASNet scoring stage uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/gemm_epi_bench scripts/dev/gemm_epi_bench.cu \
//        -Isummarizer_b200/csrc -Lsummarizer_b200 -lsummarizer_b200 && LD_LIBRARY_PATH=summarizer_b200 /tmp/gemm_epi_bench
#include <stdio.h>
#include <vector>
#include <cuda_bf16.h>
#include "smz_gemm.cuh"

using smz::GemmEpilogue; using smz::GemmProblem;
typedef __nv_bfloat16 bf16;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void fill(bf16 *p, size_t n, float s) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)(i * 2654435761u) ^ (unsigned)(i >> 7);
        p[i] = __float2bfloat16(s * (((h >> 8) & 0xffff) / 65536.f - 0.5f));
    }
}
__global__ void fillf(float *p, size_t n, float v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

template <class F> float time_it(F f, int iters = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < iters; i++) f();
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters * 1000.f;
}

int main() {
    const int R = 32000, D = 1024, T = 2000, NV = 16, LD = 2048;
    bf16 *x, *w, *qkv, *P, *o; float *f32, *stat, *vec;
    CK(cudaMalloc(&x, (size_t)R * D * 2)); CK(cudaMalloc(&w, (size_t)3 * D * D * 2)); CK(cudaMalloc(&qkv, (size_t)R * 3 * D * 2));
    CK(cudaMalloc(&P, (size_t)R * LD * 2)); CK(cudaMalloc(&o, (size_t)R * D * 2)); CK(cudaMalloc(&f32, (size_t)R * LD * 4));
    CK(cudaMalloc(&stat, (size_t)R * 16 * 3 * 4 + 64)); CK(cudaMalloc(&vec, 4 * 3072));
    fill<<<1024, 256>>>(x, (size_t)R * D, 1.f); fill<<<1024, 256>>>(w, (size_t)3 * D * D, 0.05f);
    fill<<<1024, 256>>>(qkv, (size_t)R * 3 * D, 0.3f); fill<<<1024, 256>>>(P, (size_t)R * LD, 1.f);
    fillf<<<64, 256>>>(vec, 3072, 0.01f); fillf<<<1024, 256>>>(stat, (size_t)R * 16 * 3 + 16, 1.f);
    CK(cudaDeviceSynchronize());
    auto dense = [&](int M, int N, int K, int ldc, int ldr) { GemmProblem g = {}; g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.ldr = ldr; g.tiles_n = (N + 255) / 256; return g; };
    auto report = [&](const char *name, double flop, float us) { printf("%-44s %8.1f us  %7.1f TFLOP/s\n", name, us, flop / us / 1e6); fflush(stdout); };
    cudaStream_t st = 0;
    int rc = 0;
    // ---- dense shapes
    report("qkv   N=3072 plain bf16", 2.0 * R * 3072 * D, time_it([&] { rc |= smz::gemm_bf16_tn(x, R, D, D, w, 3072, D, D, nullptr, 1, smz::gemm_tiles(R, 3072), dense(R, 3072, D, 3072, 0), GemmEpilogue{qkv, nullptr, nullptr, 1.f, 0}, st); }));
    report("      N=1024 plain bf16", 2.0 * R * D * D, time_it([&] { rc |= smz::gemm_bf16_tn(x, R, D, D, w, D, D, D, nullptr, 1, smz::gemm_tiles(R, D), dense(R, D, D, D, 0), GemmEpilogue{o, nullptr, nullptr, 1.f, 0}, st); }));
    report("      N=1024 bf16 + bias", 2.0 * R * D * D, time_it([&] { rc |= smz::gemm_bf16_tn(x, R, D, D, w, D, D, D, nullptr, 1, smz::gemm_tiles(R, D), dense(R, D, D, D, 0), GemmEpilogue{o, vec, nullptr, 1.f, 0}, st); }));
    report("out   N=1024 bf16 + res bf16", 2.0 * R * D * D, time_it([&] { rc |= smz::gemm_bf16_tn(x, R, D, D, w, D, D, D, nullptr, 1, smz::gemm_tiles(R, D), dense(R, D, D, D, D), GemmEpilogue{o, nullptr, qkv, 1.f, 0}, st); }));
    report("out   N=1024 bf16 + res bf16 + LN_STATS", 2.0 * R * D * D, time_it([&] { GemmEpilogue e{o, nullptr, qkv, 1.f, smz::GEMM_LN_STATS}; e.stat_out = stat; rc |= smz::gemm_bf16_tn(x, R, D, D, w, D, D, D, nullptr, 1, smz::gemm_tiles(R, D), dense(R, D, D, D, D), e, st); }));
    report("out   N=1024 f32 + res bf16 (old)", 2.0 * R * D * D, time_it([&] { rc |= smz::gemm_bf16_tn(x, R, D, D, w, D, D, D, nullptr, 1, smz::gemm_tiles(R, D), dense(R, D, D, D, D), GemmEpilogue{f32, nullptr, qkv, 1.f, smz::GEMM_OUT_F32}, st); }));
    report("k1    N=1024 head (old)", 2.0 * R * D * D, time_it([&] { GemmEpilogue e{f32, vec, nullptr, 1.f, smz::GEMM_RELU | smz::GEMM_OUT_F32 | smz::GEMM_ROWSTATS | smz::GEMM_NO_STORE}; e.stat_w = vec; e.stat_out = stat; rc |= smz::gemm_bf16_tn(x, R, D, D, w, D, D, D, nullptr, 1, smz::gemm_tiles(R, D), dense(R, D, D, D, 0), e, st); }));
    report("k1    N=1024 head + LN_FOLD", 2.0 * R * D * D, time_it([&] { GemmEpilogue e{f32, vec, nullptr, 1.f, smz::GEMM_RELU | smz::GEMM_OUT_F32 | smz::GEMM_ROWSTATS | smz::GEMM_NO_STORE | smz::GEMM_LN_FOLD}; e.stat_w = vec; e.stat_out = stat; e.ln_stats = stat; e.ln_c = vec; e.ln_slots = 8; e.ln_width = 1024; e.ln_eps = 1e-6f; rc |= smz::gemm_bf16_tn(x, R, D, D, w, D, D, D, nullptr, 1, smz::gemm_tiles(R, D), dense(R, D, D, D, 0), e, st); }));
    // ---- attention shapes: 16 videos of T = 2000
    std::vector<GemmProblem> pr(2 * NV);
    int tl = 0, tp = 0;
    for (int v = 0; v < NV; v++) {
        GemmProblem &a = pr[v]; a = GemmProblem{};
        a.a_row0 = v * T; a.a_col0 = 0; a.b_row0 = v * T; a.b_col0 = D; a.M = T; a.N = T; a.K = D; a.tile0 = tl; a.c_off = (int64_t)v * T * LD; a.ldc = LD;
        a.tiles_n = (T + 255) / 256; a.r_off = v * T; tl += smz::gemm_tiles(T, T);
        GemmProblem &b = pr[NV + v]; b = GemmProblem{};
        b.a_row0 = v * T; b.a_col0 = 0; b.b_row0 = v * T; b.b_col0 = 2 * D; b.M = T; b.N = D; b.K = T; b.tile0 = tp; b.c_off = (int64_t)v * T * D; b.ldc = D;
        b.tiles_n = D / 256; b.r_off = v * T; tp += smz::gemm_tiles(T, D);
    }
    GemmProblem *dpr; CK(cudaMalloc(&dpr, pr.size() * sizeof(GemmProblem))); CK(cudaMemcpy(dpr, pr.data(), pr.size() * sizeof(GemmProblem), cudaMemcpyHostToDevice));
    const double fl_att = 2.0 * NV * (double)T * T * D;
    report("logits 16x2000 plain bf16", fl_att, time_it([&] { rc |= smz::gemm_bf16_tn(qkv, R, 3 * D, 3 * D, qkv, R, 3 * D, 3 * D, dpr, NV, tl, GemmProblem{}, GemmEpilogue{P, nullptr, nullptr, 0.06f, 0}, st); }));
    report("logits 16x2000 f32", fl_att, time_it([&] { rc |= smz::gemm_bf16_tn(qkv, R, 3 * D, 3 * D, qkv, R, 3 * D, 3 * D, dpr, NV, tl, GemmProblem{}, GemmEpilogue{f32, nullptr, nullptr, 0.06f, smz::GEMM_OUT_F32}, st); }));
    int *guard = reinterpret_cast<int *>(stat + (size_t)R * 16 * 3);
    report("logits 16x2000 EXP + row sums", fl_att, time_it([&] { GemmEpilogue e{P, nullptr, nullptr, 0.06f, smz::GEMM_EXP | smz::GEMM_ROWSTATS}; e.stat_out = stat; e.stat_slots = 16; e.guard = guard; rc |= smz::gemm_bf16_tn(qkv, R, 3 * D, 3 * D, qkv, R, 3 * D, 3 * D, dpr, NV, tl, GemmProblem{}, e, st); }));
    report("pv     16x2000 MN-major V plain", fl_att, time_it([&] { rc |= smz::gemm_bf16(false, true, P, R, LD, LD, qkv, R, 3 * D, 3 * D, dpr + NV, NV, tp, GemmProblem{}, GemmEpilogue{o, nullptr, nullptr, 1.f, 0}, st); }));
    report("pv     16x2000 MN-major V SCALE_STATS", fl_att, time_it([&] { GemmEpilogue e{o, stat, nullptr, 1.f, smz::GEMM_SCALE_STATS}; e.stat_slots = 16; rc |= smz::gemm_bf16(false, true, P, R, LD, LD, qkv, R, 3 * D, 3 * D, dpr + NV, NV, tp, GemmProblem{}, e, st); }));
    // ---- the folded fast path (smz_vasnet.cu, fast_chunk)
    float *lnstat; CK(cudaMalloc(&lnstat, (size_t)R * 8 * 3 * 4));
    report("proj  N=2048 plain bf16", 2.0 * R * 2048 * D, time_it([&] { rc |= smz::gemm_bf16_tn(x, R, D, D, w, 2048, D, D, nullptr, 1, smz::gemm_tiles(R, 2048), dense(R, 2048, D, 3072, 0), GemmEpilogue{qkv, nullptr, nullptr, 1.f, 0}, st); }));
    report("pv'   SCALE_STATS + res + LN_STATS + f16", fl_att, time_it([&] { GemmEpilogue e{o, stat, x, 1.f, smz::GEMM_SCALE_STATS | smz::GEMM_RES_AT_C | smz::GEMM_LN_STATS | smz::GEMM_OUT_F16}; e.scale_slots = 16; e.stat_out = lnstat; e.stat_slots = 8; e.guard = guard; rc |= smz::gemm_bf16(false, true, P, R, LD, LD, qkv, R, 3 * D, 3 * D, dpr + NV, NV, tp, GemmProblem{}, e, st); }));
    report("k1    head + LN_FOLD (f16 operands)", 2.0 * R * D * D, time_it([&] { GemmEpilogue e{f32, vec, nullptr, 1.f, smz::GEMM_RELU | smz::GEMM_OUT_F32 | smz::GEMM_ROWSTATS | smz::GEMM_NO_STORE | smz::GEMM_LN_FOLD | smz::GEMM_A_F16 | smz::GEMM_B_F16}; e.stat_w = vec; e.stat_out = stat; e.ln_stats = lnstat; e.ln_c = vec; e.ln_slots = 8; e.ln_width = 1024; e.ln_eps = 1e-6f; rc |= smz::gemm_bf16_tn(x, R, D, D, w, D, D, D, nullptr, 1, smz::gemm_tiles(R, D), dense(R, D, D, D, 0), e, st); }));
    printf("rc %d %s\n", rc, rc ? smz_last_error() : "");
    return rc;
}
