// Annotator summaries as 1 bit per frame: the host-to-device staging form of utils/eval.py:125-165's
// `user_summary` input.  evaluate_summary binarises it first (`user_summary[user_summary > 0] = 1`, eval.py:148-149),
// so (x > 0) is all the F-score ever reads: packing it on the host before the PCIe copy moves 32x fewer bytes than
// the float32 rows (2.4 MB -> 75 KB per sweep video), and the device side becomes a popcount over words.
//   smz_host_pack_user_summary  host threads, SSE2 compare + movemask (runs on the CPU by design: it is the copy's
//                               staging step, not a compute fallback — no result is produced here)
//   smz_fscore_packed           device: overlap / annotator counts from the packed rows, then the same
//                               fscore_final arithmetic as smz_fscore (bit-identical F).
#include <emmintrin.h>

#include <stdlib.h>

#include <thread>
#include <vector>

#include "smz_common.cuh"
#include "smz_tc.cuh"

namespace smz {
int launch_fscore_final(const smz_video_desc *desc, int n_videos, const int32_t *msum, const int32_t *overlap,
                        const int32_t *gsum, float *f, double *avg_f, double *max_f, cudaStream_t st);
}

namespace {

void pack_row(const float *src, int n, uint32_t *dst) {
    const __m128 zero = _mm_setzero_ps();
    const int full = n / 32;
    for (int w = 0; w < full; w++) {
        const float *p = src + 32 * w;
        uint32_t bits = 0;
        for (int q = 0; q < 8; q++)
            bits |= (uint32_t)_mm_movemask_ps(_mm_cmpgt_ps(_mm_loadu_ps(p + 4 * q), zero)) << (4 * q);
        dst[w] = bits;
    }
    if (n % 32) {
        uint32_t bits = 0;
        for (int j = 32 * full; j < n; j++) bits |= (src[j] > 0.f ? 1u : 0u) << (j & 31);
        dst[full] = bits;
    }
}

__global__ void __launch_bounds__(256)
fscore_bits_kernel(const smz_video_desc *__restrict__ desc, int v0, const uint32_t *__restrict__ ubits,
                   const int64_t *__restrict__ bits_off, const uint32_t *__restrict__ mask,
                   int32_t *__restrict__ overlap, int32_t *__restrict__ gsum) {
    const int v = v0 + blockIdx.x;
    const smz_video_desc d = desc[v];
    const int W = (d.n_frames + 31) >> 5;
    const uint32_t *vm = mask + d.mask_off;
    const uint32_t *ub = ubits + bits_off[v];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int u = warp; u < d.n_users; u += 8) {
        const uint32_t *row = ub + (int64_t)u * W;
        int ov = 0, gs = 0;
        for (int w = lane; w < W; w += 32) {
            const uint32_t x = __ldg(row + w);
            ov += __popc(x & __ldg(vm + w));
            gs += __popc(x);
        }
        ov = __reduce_add_sync(0xffffffffu, ov);
        gs = __reduce_add_sync(0xffffffffu, gs);
        if (lane == 0) { overlap[d.ucount_off + u] = ov; gsum[d.ucount_off + u] = gs; }
    }
}

// Device-side packing: float32 annotator rows -> 1 bit per frame, the layout smz_host_pack_user_summary produces.
// This is the HBM-bound half of evaluate_summary (4 * n_users * n_frames bytes per video) and it does not depend on the
// machine summary: a caller that evaluates several score sets against the same annotations packs once and uses
// smz_fscore_packed / smz_eval_batch(user_bits) afterwards (1/32 of the bytes per evaluation).  Measured 0.97 of the HBM
// copy bandwidth.  (Running it beside the knapsack kernels inside one smz_eval_batch call was tried in round 2 and lost
// to the fused tail: see DESIGN.md.)
// The kernel saturates HBM from a SMALL footprint: one CTA of 4 warps per SM, and the bytes in flight live in shared
// memory, not in registers — every warp keeps a ring of PK_STAGES
// bulk copies (cp.async.bulk global -> shared, 4 KB = 1024 frames of one row each, completion on an mbarrier, L2
// evict-first: the rows are read once), 128 KB in flight per SM.  A warp turns one 4 KB chunk into 32 words: lane l owns
// frames 32 l .. 32 l + 31 = word l of the chunk; it reads its eight float4 in the rotated order (q + l) % 8 so that the
// eight lanes of a shared-memory wavefront hit eight different bank groups, ORs the nibbles into its word and the warp
// stores 128 contiguous bytes.
constexpr int PK_WARPS = 4, PK_STAGES = 8, PK_CHUNK = 4096;
constexpr int PK_SMEM = PK_WARPS * PK_STAGES * PK_CHUNK + PK_WARPS * PK_STAGES * 8 + 128;

__device__ __forceinline__ uint32_t pos_flag(float x) { return (uint32_t)__vimin_s32_relu(__float_as_int(x), 1); }   // x > 0 (see smz_fscore_dev.cuh)

__device__ __forceinline__ void bulk_load(void *smem_dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smztc::smem_u32(smem_dst)), "l"(src), "r"(bytes), "r"(smztc::smem_u32(bar)), "l"(policy) : "memory");
}

// one float4 of flags -> a nibble at bit position 4 * slot
__device__ __forceinline__ uint32_t nibble_at(const float4 x, int slot) {
    return ((pos_flag(x.x) + 2u * pos_flag(x.y)) + 4u * (pos_flag(x.z) + 2u * pos_flag(x.w))) << (4 * slot);
}

__global__ void __launch_bounds__(PK_WARPS * 32)
pack_user_bits_kernel(const smz_video_desc *__restrict__ desc, int n_videos, const float *__restrict__ user,
                      const int64_t *__restrict__ bits_off, uint32_t *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t pk_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *ring = pk_smem + warp * PK_STAGES * PK_CHUNK;
    uint64_t *bar = reinterpret_cast<uint64_t *>(pk_smem + PK_WARPS * PK_STAGES * PK_CHUNK) + warp * PK_STAGES;
    if (lane == 0) {
        for (int s = 0; s < PK_STAGES; s++) smztc::mbar_init(&bar[s], 1);
        smztc::fence_mbar_init();
    }
    __syncwarp();
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    uint32_t it = 0;                 // chunks this warp has consumed: stage = it % PK_STAGES, parity = (it / PK_STAGES) & 1
    for (int v = blockIdx.x; v < n_videos; v += gridDim.x) {
        const smz_video_desc d = desc[v];
        const int n_frames = d.n_frames, W = (n_frames + 31) >> 5;
        const int trips = (n_frames + 1023) >> 10;
        const int n_tasks = d.n_users * trips;
        uint32_t *vb = out + bits_off[v];
        const bool vec = (((d.user_off | d.user_ld) & 3) == 0) && ((reinterpret_cast<uintptr_t>(user) & 15) == 0);
        if (!vec) {                  // rows that are not 16-byte aligned: plain loads
            for (int task = warp; task < n_tasks; task += PK_WARPS) {
                const int u = task / trips, t = task - u * trips;
                const float *row = user + d.user_off + (int64_t)u * d.user_ld;
                const int f = (t << 10) + lane * 32;
                uint32_t word = 0u;
#pragma unroll 1
                for (int j = 0; j < 32; j++)
                    if (f + j < n_frames && row[f + j] > 0.f) word |= 1u << j;
                const int widx = (t << 5) + lane;
                if (widx < W) vb[(int64_t)u * W + widx] = word;
            }
            continue;
        }
        auto issue = [&](int task, uint32_t slot) {      // lane 0: chunk `task` of this video -> ring stage slot % PK_STAGES
            const int u = task / trips, t = task - u * trips;
            const float *src = user + d.user_off + (int64_t)u * d.user_ld + (t << 10);
            const int left = (int)d.user_ld - (t << 10);                 // floats up to the end of the padded row (multiple of 4)
            const uint32_t bytes = (uint32_t)(left < 1024 ? left : 1024) * 4u;
            const uint32_t s = slot % PK_STAGES;
            smztc::mbar_arrive_expect_tx(&bar[s], bytes);
            bulk_load(ring + s * PK_CHUNK, src, bytes, &bar[s], policy);
        };
        if (lane == 0)
            for (int k = 0; k < PK_STAGES; k++)
                if (warp + k * PK_WARPS < n_tasks) issue(warp + k * PK_WARPS, it + k);
        for (int task = warp; task < n_tasks; task += PK_WARPS, ++it) {
            const uint32_t s = it % PK_STAGES;
            smztc::mbar_wait(&bar[s], (it / PK_STAGES) & 1u);
            const float4 *buf = reinterpret_cast<const float4 *>(ring + s * PK_CHUNK) + lane * 8;
            float4 x[8];
#pragma unroll
            for (int q = 0; q < 8; q++) x[q] = buf[(q + lane) & 7];
            const int u = task / trips, t = task - u * trips;
            uint32_t word = 0u;
#pragma unroll
            for (int q = 0; q < 8; q++) word |= nibble_at(x[q], (q + lane) & 7);
            // every lane has CONSUMED its data (a load merely issued could still be in flight when the refill lands):
            // only now may the stage be handed back to the copy engine
            __syncwarp();
            if (lane == 0 && task + PK_STAGES * PK_WARPS < n_tasks) issue(task + PK_STAGES * PK_WARPS, it + PK_STAGES);
            const int left = n_frames - ((t << 10) + lane * 32);    // frames past the end (row padding, stale ring bytes) count as 0
            if (left < 32) word &= left > 0 ? (1u << left) - 1u : 0u;
            const int widx = (t << 5) + lane;
            if (widx < W) vb[(int64_t)u * W + widx] = word;
        }
    }
}

}  // namespace

namespace smz {
int launch_pack_user_bits(const smz_video_desc *desc, int n_videos, const float *user, const int64_t *bits_off,
                          uint32_t *out, cudaStream_t st) {
    if (n_videos <= 0) return SMZ_OK;
    static int ctas_per_sm = -1;
    if (ctas_per_sm < 0) { const char *e = getenv("SMZ_PACK_CTAS_PER_SM"); ctas_per_sm = (e != nullptr && atoi(e) > 0) ? atoi(e) : 1; }
    static bool attr_set[64] = {false};
    int dev = 0;
    SMZ_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        SMZ_CUDA_CHECK(cudaFuncSetAttribute((const void *)pack_user_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM));
        attr_set[dev] = true;
    }
    int grid = sm_count() * ctas_per_sm;
    if (grid > n_videos) grid = n_videos;
    pack_user_bits_kernel<<<grid, PK_WARPS * 32, PK_SMEM, st>>>(desc, n_videos, user, bits_off, out);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}

int launch_fscore_bits(const smz_video_desc *desc, int n_videos, const uint32_t *user_bits, const int64_t *bits_off,
                       const uint32_t *mask, int32_t *overlap, int32_t *gsum, cudaStream_t st) {
    fscore_bits_kernel<<<n_videos, 256, 0, st>>>(desc, 0, user_bits, bits_off, mask, overlap, gsum);
    SMZ_CUDA_CHECK(cudaGetLastError());
    return SMZ_OK;
}
}  // namespace smz

// h_desc / h_user / h_bits_off / h_bits are HOST pointers.  Row u of video v: h_user + user_off + u*user_ld
// (n_frames floats) -> h_bits + h_bits_off[v] + u*ceil(n_frames/32) words, bit j of word w = frame 32w+j > 0.
extern "C" int smz_host_pack_user_summary(const smz_video_desc *h_desc, int n_videos, const float *h_user,
                                          const int64_t *h_bits_off, uint32_t *h_bits, int n_threads) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0 && h_desc && h_user && h_bits_off && h_bits, "host_pack_user_summary: bad argument");
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_videos) n_threads = n_videos;
    auto work = [=](int t) {
        for (int v = t; v < n_videos; v += n_threads) {
            const smz_video_desc &d = h_desc[v];
            const int W = (d.n_frames + 31) >> 5;
            for (int u = 0; u < d.n_users; u++)
                pack_row(h_user + d.user_off + (int64_t)u * d.user_ld, d.n_frames, h_bits + h_bits_off[v] + (int64_t)u * W);
        }
    };
    if (n_threads == 1) { work(0); return SMZ_OK; }
    std::vector<std::thread> pool;
    pool.reserve(n_threads);
    for (int t = 0; t < n_threads; t++) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
    return SMZ_OK;
}

extern "C" int smz_fscore_packed(const smz_video_desc *desc, int n_videos, const uint32_t *user_bits, const int64_t *bits_off,
                                 const uint32_t *mask, const int32_t *msum, int32_t *overlap, int32_t *gsum, float *f,
                                 double *avg_f, double *max_f, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0, "negative size");
    SMZ_REQUIRE(desc && user_bits && bits_off && mask && msum && overlap && gsum && f, "NULL pointer");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = smz::launch_fscore_bits(desc, n_videos, user_bits, bits_off, mask, overlap, gsum, st);
    if (rc != SMZ_OK) return rc;
    return smz::launch_fscore_final(desc, n_videos, msum, overlap, gsum, f, avg_f, max_f, st);
}

// DEVICE: the same packing as smz_host_pack_user_summary, from device-resident float32 rows (user_summary laid out by
// desc.user_off / user_ld) into user_bits + bits_off[v] (device array of per-video word offsets).
extern "C" int smz_pack_user_bits(const smz_video_desc *desc, int n_videos, const float *user_summary, const int64_t *bits_off,
                                  uint32_t *user_bits, void *stream) {
    if (n_videos == 0) return SMZ_OK;
    SMZ_REQUIRE(n_videos > 0 && desc && user_summary && bits_off && user_bits, "pack_user_bits: bad argument");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    return smz::launch_pack_user_bits(desc, n_videos, user_summary, bits_off, user_bits, (cudaStream_t)stream);
}
