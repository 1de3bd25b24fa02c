// VASNet scorer (models/vasnet.py:92-148) on the tcgen05 GEMM building block.
//
// Two forward paths over the packed batch (rows of all videos, [sum T, 1024]), chunked by rows:
//
// (1) the literal chain — training (the whole batch is one chunk, every intermediate stays in the work buffer for
//     smz_vasnet_backward) and the wide-range inference path (no status word; all bf16 / fp32):
//   xb  = bf16(x)                                              (skipped when the features are bf16)
//   QK  = xb . [Wq;Wk]^T                      [R, 2048] bf16    vasnet.py:114-115, one packed GEMM
//   Vt  = Wv . xb^T                           [1024, R] bf16    vasnet.py:116, produced transposed so that
//                                                               alpha.V is a K-major GEMM too (training; inference
//                                                               packs Q|K|V into ONE [R, 3072] GEMM and alpha.V reads
//                                                               V in place as an MN-major B operand)
//   per attention sub-chunk:
//     S   = scale * Q_v . K_v^T               [T, T]   fp32     vasnet.py:118-119, one GEMM problem per video
//     P   = dropout(softmax(mask(S)))         [T, ld]  bf16     vasnet.py:121-130 (zero padded columns)
//     O   = P . V_v                           [R, 1024] bf16    vasnet.py:131
//   Y   = O . Wo^T + x                        [R, 1024] fp32    vasnet.py:132-135
//   Yn  = LayerNorm(dropout(Y))               bf16              vasnet.py:136-137
//   H   = relu(Yn . W1^T + b1)                fp32              vasnet.py:140-141  (inference: never stored, head epilogue)
//   s   = sigmoid(LayerNorm(dropout(H)) . w2 + b2)              vasnet.py:142-145
//
// (2) the fast inference path (fast_chunk below): the same function with its linear maps folded and every row-wise
//     step in a GEMM epilogue — [V'|G] projection, exp-logits, alpha.V' (+ scale, residual, LayerNorm sums), k1 (+ folded
//     LayerNorm, head sums), one per-row kernel; float16 where the range can be checked, a status word instead of a
//     device-side fallback.
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "smz_gemm.cuh"
#include "smz_rows.cuh"

namespace {

using smz::GemmEpilogue;
using smz::GemmProblem;
using smz::kFeat;
typedef __nv_bfloat16 bf16;

// Chunk sizes trade L2 residency of the intermediates against wave quantisation of the per-video
// attention problems (one T=2000 video is only 64 logits tiles / 32 alpha.V tiles for 74 CTA pairs): a sub-chunk
// of 16 sweep videos gives 512 alpha.V tiles = 6.9 waves (8 videos: 3.5 waves, 13 % of the launch idle in the tail;
// measured 1.5 % of the sweep).
constexpr int kRowChunk = 32768;                // rows per chunk of the row-wise GEMMs (1000+ tiles: < 4 % wave tail)
// logits in flight per sub-chunk (elements).  Development knob: SMZ_VASNET_LOGIT_MELEMS overrides (in Mi elements).
int64_t logit_budget() {
    static int64_t v = -1;
    if (v < 0) { const char *e = getenv("SMZ_VASNET_LOGIT_MELEMS"); v = (e != nullptr && atoi(e) > 0) ? ((int64_t)atoi(e) << 20) : (68ll << 20); }
    return v;
}

inline int64_t up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct Sub { int v0, v1, rows, ld; };
// In inference mode the buffers are recycled chunk after chunk (all offsets 0); in training mode every
// video is its own chunk and owns a region of each buffer (rows at row0, V^T / logits at vt_off / lg_off).
struct Chunk { int v0, v1, row0, rows; int64_t vt_off, lg_off; std::vector<Sub> subs; };

struct Plan {
    std::vector<Chunk> chunks;
    int n_videos = 0, total_rows = 0, max_rows = 0;
    int64_t max_logits = 0;
    bool training = false, x_bf16 = false;
    int64_t rows_cap = 0;   // row capacity of the row-wise buffers (and stride of the stats arrays)
    int lanes = 1;          // inference with several chunks: two copies of the recycled buffers, one per stream
    int64_t lane_bytes = 0;
    int64_t off_xb, off_qk, off_vt, off_o, off_y, off_yn, off_h, off_s, off_p, off_alpha, off_hstat, off_sstat, off_lnstat, off_probs, off_dropoff,
        off_stats, total;
    // backward-only buffers (training)
    int64_t off_dh, off_dyn, off_dy, off_dyf, off_do, off_dqk, off_dvt, off_dp, off_ds;
};

int make_plan(const int32_t *cu, int n_videos, bool training, bool x_bf16, Plan *pl) {
    SMZ_REQUIRE(cu != nullptr && n_videos > 0, "vasnet: empty batch");
    SMZ_REQUIRE(cu[0] == 0, "vasnet: cu_seqlens[0] must be 0");
    for (int v = 0; v < n_videos; v++) SMZ_REQUIRE(cu[v + 1] > cu[v], "vasnet: video %d has no frames", v);
    pl->n_videos = n_videos;
    pl->total_rows = cu[n_videos];
    pl->training = training;
    pl->x_bf16 = x_bf16;
    int v = 0;
    int64_t vt_elems = 0, lg_elems = 0;
    while (v < n_videos) {
        Chunk c;
        c.v0 = v; c.row0 = cu[v];
        int rows = 0;
        while (v < n_videos && (rows == 0 || (!training && rows + (cu[v + 1] - cu[v]) <= kRowChunk))) { rows += cu[v + 1] - cu[v]; ++v; }
        c.v1 = v; c.rows = rows;
        c.vt_off = training ? vt_elems : 0;
        c.lg_off = training ? lg_elems : 0;
        int64_t chunk_logits = 0;
        int u = c.v0;
        while (u < c.v1) {
            Sub s;
            s.v0 = u; s.rows = 0;
            int maxT = 0;
            while (u < c.v1) {
                const int T = cu[u + 1] - cu[u];
                const int nm = T > maxT ? T : maxT;
                const int64_t elems = (int64_t)(s.rows + T) * up(nm + 7, 64);
                if (s.rows > 0 && elems > logit_budget()) break;
                maxT = nm; s.rows += T; ++u;
            }
            s.v1 = u; s.ld = (int)up(maxT + 7, 64);   // + up to 7 leading pad columns, see build_problems
            const int64_t elems = (int64_t)s.rows * s.ld;
            if (elems > pl->max_logits) pl->max_logits = elems;
            chunk_logits += elems;
            c.subs.push_back(s);
        }
        if (training) { vt_elems += (int64_t)kFeat * up(rows, 8); lg_elems += chunk_logits; }
        if (c.rows > pl->max_rows) pl->max_rows = c.rows;
        pl->chunks.push_back(c);
    }
    const int64_t R = training ? up(pl->total_rows, 8) : up(pl->max_rows, 8);
    const int64_t VT = training ? vt_elems : (int64_t)kFeat * R;
    const int64_t LG = training ? lg_elems : pl->max_logits;
    pl->rows_cap = R;
    int64_t o = 0;
    auto take = [&](int64_t bytes) { const int64_t at = o; o += up(bytes, 1024); return at; };
    pl->off_xb = take(x_bf16 ? 0 : R * kFeat * 2);
    pl->off_qk = take(R * (training ? 2 : 3) * kFeat * 2);      // inference: Q | K | V packed, [R, 3072]
    pl->off_vt = take(training ? VT * 2 : 0);
    pl->off_o = take(R * kFeat * 2);
    pl->off_y = take(R * kFeat * 4);
    pl->off_yn = take(R * kFeat * 2);
    pl->off_h = take(R * kFeat * 4);
    pl->off_s = take(LG * 4);
    pl->off_p = take(LG * 2);
    pl->off_alpha = take(training ? LG * 2 : 0);
    pl->off_hstat = take(training ? 0 : R * (2 * kFeat / smz::GEMM_BN) * 3 * 4);   // fused head: [R][8 slots][3]
    {   // fused-exp attention (inference): row-sum slots [chunk rows][2 * ceil(ld / 256)][3]
        int64_t slots = 0;
        for (const Chunk &c : pl->chunks)
            for (const Sub &s : c.subs) { const int64_t k = 2 * ((s.ld + 255) / 256); if (k > slots) slots = k; }
        pl->off_sstat = take(training ? 0 : R * slots * 12 + 64);
    }
    pl->off_lnstat = take(training ? 0 : R * (2 * kFeat / smz::GEMM_BN) * 3 * 4);  // LayerNorm-1 sums of the out-proj epilogue
    pl->lane_bytes = o;
    pl->lanes = (!training && pl->chunks.size() > 1) ? 2 : 1;
    if (pl->lanes == 2) o += pl->lane_bytes;     // second copy of everything above
    pl->off_probs = take((int64_t)n_videos * 2 * sizeof(GemmProblem));
    pl->off_dropoff = take(training ? (int64_t)n_videos * 8 : 0);
    pl->off_stats = take(training ? R * 4 * 4 : 0);
    const int64_t t = training ? 1 : 0;
    pl->off_dh = take(t * R * kFeat * 2);
    pl->off_dyn = take(t * R * kFeat * 4);
    pl->off_dy = take(t * R * kFeat * 2);
    pl->off_dyf = take(t * R * kFeat * 4);
    pl->off_do = take(t * R * kFeat * 2);
    pl->off_dqk = take(t * R * 2 * kFeat * 2);
    pl->off_dvt = take(t * VT * 2);
    pl->off_dp = pl->off_s;              // dP (float32) reuses the logits buffer, dS (bf16) gets its own
    pl->off_ds = take(t * LG * 2);
    pl->total = o;
    return SMZ_OK;
}

// host-side GEMM problem tables: [0, n) logits problems, [n, 2n) alpha.V problems
// fast (inference with folded weights): the packed projection rows are [V' | G] (V' = x.(Wo Wv)^T in columns [0, 1024),
// G = x.(Wq^T Wk) in [1024, 2048)); logits = G . x^T reads the features as its B operand; alpha.V' reads V' in place
// (MN-major) and r_off carries the chunk-local row of the video for the row statistics (the residual sits at c_off).
void build_problems(const Plan &pl, const int32_t *cu, bool fast, std::vector<GemmProblem> *out) {
    const bool packed_v = !pl.training;       // inference: V lives in columns [2048, 3072) of the packed Q|K|V rows
    const int n = pl.n_videos;
    out->assign((size_t)2 * n, GemmProblem{});
    for (const Chunk &c : pl.chunks)
        for (const Sub &s : c.subs) {
            int tile_s = 0, tile_pv = 0, sub_row = 0;
            for (int v = s.v0; v < s.v1; v++) {
                const int T = cu[v + 1] - cu[v];
                const int crow = cu[v] - c.row0;          // chunk-local row of the video
                // TMA needs 16-byte aligned coordinates along the contiguous dimension: the video's columns
                // of V^T start at crow, so alpha.V reads V^T from crow - lead and P carries `lead` zero
                // columns in front (written by the softmax kernel).
                // Inference reads V [T, 1024] in place as an MN-major B operand (k = frame = TMA row coordinate, no
                // alignment constraint): no lead there.
                const int lead = packed_v ? 0 : (crow & 7);
                GemmProblem &a = (*out)[v];
                a.a_row0 = crow; a.a_col0 = fast ? kFeat : 0; a.b_row0 = crow; a.b_col0 = fast ? 0 : kFeat;
                a.M = T; a.N = T; a.K = kFeat; a.tile0 = tile_s;
                a.c_off = (int64_t)sub_row * s.ld; a.ldc = s.ld;
                a.tiles_n = (T + smz::GEMM_BN - 1) / smz::GEMM_BN;
                a.pad = lead;
                a.r_off = fast ? crow : sub_row;        // GEMM_ROWSTATS slots: row of the video inside the chunk / sub-chunk
                tile_s += smz::gemm_tiles(T, T);
                GemmProblem &b = (*out)[n + v];
                b.a_row0 = sub_row; b.a_col0 = 0; b.b_row0 = 0; b.b_col0 = crow - lead;
                if (packed_v) { b.b_row0 = crow; b.b_col0 = fast ? 0 : 2 * kFeat; }
                b.M = T; b.N = kFeat; b.K = T + lead; b.tile0 = tile_pv;
                b.c_off = (int64_t)crow * kFeat; b.ldc = kFeat;
                b.tiles_n = kFeat / smz::GEMM_BN;
                b.r_off = fast ? crow : sub_row;        // fast: row of the softmax row sums (read) and LayerNorm sums (written)
                tile_pv += smz::gemm_tiles(T, kFeat);
                sub_row += T;
            }
        }
}

// Two internal streams per (device, caller stream): consecutive chunks run on alternating streams (each with its own
// copy of the chunk buffers) so that the tail wave of one chunk's GEMM overlaps the next chunk's kernels; the training
// path forks its independent GEMMs onto them.  They are forked from / joined back into the caller's stream with events,
// so the call stays asynchronous and ordered.  Keyed by the CALLER'S STREAM, not by host thread: fold-concurrent
// training drives one stream per fold, and torch runs every fold's backward pass on the same autograd thread — two
// folds sharing events would tie one fold's graph capture to the other's eager work.
struct SideStreams { cudaStream_t s[2]; cudaEvent_t fork, join[2], ev[6]; bool ok; };
SideStreams *side_streams(cudaStream_t caller) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, SideStreams *> table;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    SideStreams *&slot = table[std::make_pair(dev, caller)];
    if (slot == nullptr) {
        SideStreams *t = new SideStreams();
        for (int i = 0; i < 2; i++) {
            if (cudaStreamCreateWithFlags(&t->s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&t->join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&t->fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        for (int i = 0; i < 6; i++)
            if (cudaEventCreateWithFlags(&t->ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        t->ok = true;
        slot = t;
    }
    return slot;
}

GemmProblem dense_problem(int M, int N, int K, int ldc, int ldr) {
    GemmProblem g = {};
    g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.ldr = ldr;
    g.tiles_n = (N + smz::GEMM_BN - 1) / smz::GEMM_BN;
    return g;
}

// ---------------------------------------------------------------------------------------------------
// FAST inference path: the reference's forward (vasnet.py:114-145) with its linear maps folded where no non-linearity
// sits between them — exact in real arithmetic, and 22 % fewer multiply-adds than the literal chain:
//   logits_ij = (Wq x_i).(Wk x_j) = (x_i^T M) x_j          M   = Wq^T Wk     (the K projection disappears)
//   c_i = Wo sum_j a_ij Wv x_j   = sum_j a_ij (Wvo x_j)    Wvo = Wo Wv       (the output projection disappears)
//   W1 LN(y) + b1 = rstd (W1 diag(g) y - mean W1 g) + (W1 b + b1)            (the LayerNorm kernel disappears)
// and the row-wise steps fused into GEMM epilogues.  Per chunk of <= 32 768 frames:
//   [V' | G] = x . [Wvo ; M^T]^T                  one [R, 2048] GEMM
//   P        = exp(scale * G . x^T) (masked), row sums        logits epilogue, no max subtraction (range-checked)
//   y        = (P . V') / rowsum + x, row sum / sum of squares of y, float16      alpha.V' epilogue, V' read in place (MN-major)
//   score    = head(relu(rstd (W1g . y - mean c) + b1f))      k1 epilogue reduces relu(h) to three row sums + a per-row kernel
// 16-bit formats: bf16 features as given (float32 features are copied to float16 and then G / M / the logits operands
// are float16 too); V', P bf16 (alpha.V' must match P, whose un-normalised exponentials need the bf16 range); y, W1g
// float16.  tcgen05 kind::f16 wants both operands of one GEMM in the same format.  Every float16 value and every logit
// is range-checked on the fly (SMZ_VASNET_STATUS_* in *status); a violation voids the call and the caller repeats it
// on the wide-range path below (the literal chain, bf16 / fp32, max-subtracted softmax).  No launch is gated on the
// status word: an empty launch of the persistent GEMM costs ~5 us, two per sub-chunk were 4 % of the sweep.
int fast_chunk(const Plan &pl, const Chunk &c, const int32_t *cu, const smz_vasnet_params *p, const void *xb, const void *res,
               bool x_bf16, bf16 *gv, bf16 *P, bf16 *y16, float *sstat, float *lnstat, float *hstat, const GemmProblem *d_probs,
               float *scores, cudaStream_t st) {
    const int R = c.rows, n = pl.n_videos;
    const int f16 = x_bf16 ? 0 : (smz::GEMM_A_F16 | smz::GEMM_B_F16);
    int rc;
    smz::profile_mark(st, "gemm_proj");
    {
        GemmEpilogue e{gv, nullptr, nullptr, 1.f, f16 | (x_bf16 ? 0 : smz::GEMM_OUT_F16)};
        e.f16_col0 = kFeat; e.guard = p->status; e.guard_bit = SMZ_VASNET_STATUS_F16_RANGE;     // V' bf16 | G float16 (float32 features)
        rc = smz::gemm_bf16_tn(xb, R, kFeat, kFeat, x_bf16 ? p->wgv : p->wgv16, 2 * kFeat, kFeat, kFeat, nullptr, 1,
                               smz::gemm_tiles(R, 2 * kFeat), dense_problem(R, 2 * kFeat, kFeat, 2 * kFeat, 0), e, st);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "gemm_proj");
    }
    int slots = 0;
    for (const Sub &s : c.subs) { const int k = 2 * ((s.ld + 255) / 256); if (k > slots) slots = k; }
    bool ragged = false;    // a video narrower than the widest leaves row-sum slots unwritten: they must read 0
    for (int v = c.v0; v < c.v1; v++) ragged |= 2 * ((cu[v + 1] - cu[v] + 255) / 256) != slots;
    if (ragged) SMZ_CUDA_CHECK(cudaMemsetAsync(sstat, 0, (int64_t)R * slots * 3 * 4, st));
    constexpr int kLnSlots = 2 * kFeat / smz::GEMM_BN;
    for (const Sub &s : c.subs) {
        const int nv = s.v1 - s.v0;
        int tiles_s = 0, tiles_pv = 0;
        for (int v = s.v0; v < s.v1; v++) {
            const int T = cu[v + 1] - cu[v];
            tiles_s += smz::gemm_tiles(T, T);
            tiles_pv += smz::gemm_tiles(T, kFeat);
        }
        smz::profile_mark(st, "gemm_logits");
        {   // softmax is shift invariant: exp without the max subtraction is exact while |logit| <= 80 (checked)
            GemmEpilogue e{P, nullptr, nullptr, p->scale, smz::GEMM_EXP | smz::GEMM_ROWSTATS | f16};
            e.stat_out = sstat; e.stat_slots = slots; e.aperture = p->aperture; e.ignore_self = p->ignore_self;
            e.guard = p->status; e.guard_bit = SMZ_VASNET_STATUS_LOGIT_RANGE;
            rc = smz::gemm_bf16_tn(gv, R, 2 * kFeat, 2 * kFeat, xb, R, kFeat, kFeat, d_probs + s.v0, nv, tiles_s, GemmProblem{}, e, st);
            if (rc != SMZ_OK) return rc;
            SMZ_DEBUG_STEP(st, "gemm_logits_exp");
        }
        smz::profile_mark(st, "gemm_pv");
        {   // y = P . V' / rowsum + x  (float16) and the LayerNorm sums of its float32 values
            GemmEpilogue e{y16, sstat, res, 1.f, smz::GEMM_SCALE_STATS | smz::GEMM_RES_AT_C | (x_bf16 ? 0 : smz::GEMM_RES_F32) |
                                                 smz::GEMM_LN_STATS | smz::GEMM_OUT_F16};
            e.scale_slots = slots; e.stat_out = lnstat; e.stat_slots = kLnSlots;
            e.guard = p->status; e.guard_bit = SMZ_VASNET_STATUS_F16_RANGE;
            rc = smz::gemm_bf16(false, true, P, s.rows, s.ld, s.ld, gv, R, 2 * kFeat, 2 * kFeat, d_probs + n + s.v0, nv, tiles_pv,
                                GemmProblem{}, e, st);
            if (rc != SMZ_OK) return rc;
            SMZ_DEBUG_STEP(st, "gemm_pv");
        }
    }
    smz::profile_mark(st, "gemm_k1");
    {
        GemmEpilogue e{hstat, p->b1f, nullptr, 1.f, smz::GEMM_RELU | smz::GEMM_OUT_F32 | smz::GEMM_ROWSTATS | smz::GEMM_NO_STORE |
                                                   smz::GEMM_LN_FOLD | smz::GEMM_A_F16 | smz::GEMM_B_F16};
        e.stat_w = p->head_gw; e.stat_out = hstat;
        e.ln_stats = lnstat; e.ln_c = p->ln_c; e.ln_slots = kLnSlots; e.ln_width = kFeat; e.ln_eps = p->eps;
        rc = smz::gemm_bf16_tn(y16, R, kFeat, kFeat, p->w1g, kFeat, kFeat, kFeat, nullptr, 1, smz::gemm_tiles(R, kFeat),
                               dense_problem(R, kFeat, kFeat, kFeat, 0), e, st);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "gemm_k1");
    }
    smz::profile_mark(st, "head");
    rc = smz::launch_head_from_stats(hstat, kLnSlots, p->head_c, p->eps, R, scores, st);
    if (rc != SMZ_OK) return rc;
    SMZ_DEBUG_STEP(st, "head");
    smz::profile_mark(st, "");
    return SMZ_OK;
}

}  // namespace

namespace { int g_exact_softmax = 0; }
// Testing / A-B aid: on != 0 forces the exact (fp32 logits + max-subtracted softmax) attention path for inference.
extern "C" void smz_vasnet_set_exact_softmax(int on) { g_exact_softmax = on; }

extern "C" int smz_vasnet_workspace_bytes(const int32_t *h_cu_seqlens, int n_videos, int training, int x_is_bf16,
                                          int64_t *bytes) {
    SMZ_REQUIRE(bytes != nullptr, "bytes is NULL");
    Plan pl;
    SMZ_REQUIRE(!(training & SMZ_VASNET_SPLIT) || (training & 1), "vasnet: the split-bf16 mode uses the training layout (training = 1 | SMZ_VASNET_SPLIT)");
    int rc = make_plan(h_cu_seqlens, n_videos, (training & ~SMZ_VASNET_SPLIT) != 0, x_is_bf16 != 0, &pl);
    if (rc != SMZ_OK) return rc;
    *bytes = (training & SMZ_VASNET_SPLIT) ? 2 * pl.total : pl.total;     // second half: the lo planes, same offsets
    return SMZ_OK;
}

// number of kernel launches smz_vasnet_forward issues for this batch (bench.py reports it)
extern "C" int smz_vasnet_launch_count(const int32_t *h_cu_seqlens, int n_videos, int training, int x_is_bf16,
                                       int64_t *launches) {
    SMZ_REQUIRE(launches != nullptr, "launches is NULL");
    Plan pl;
    int rc = make_plan(h_cu_seqlens, n_videos, training == 1, x_is_bf16 != 0, &pl);
    if (rc != SMZ_OK) return rc;
    int64_t n = 0;
    // training (1): [cvt] QK Vt out layernorm k1 head per chunk; logits, softmax, alpha.V per sub-chunk
    // inference (0, the fast path taken with params.status): [cvt] proj k1 head per chunk; fused-exp logits, alpha.V' per sub-chunk
    // exact inference (2, no status word): [cvt] QKV out layernorm k1 head per chunk; logits, softmax, alpha.V per sub-chunk
    for (const Chunk &c : pl.chunks) {
        const int64_t subs = (int64_t)c.subs.size(), cvt = x_is_bf16 ? 0 : 1;
        n += training == 1 ? cvt + 6 + 3 * subs : (training == 2 ? cvt + 5 + 3 * subs : cvt + 3 + 2 * subs);
    }
    *launches = n;
    return SMZ_OK;
}

extern "C" int smz_vasnet_forward(const void *x, int x_is_bf16, const int32_t *h_cu_seqlens, int n_videos,
                                  const smz_vasnet_params *p, int training, const uint8_t *drop_att,
                                  const uint8_t *drop_y, const uint8_t *drop_h, float *scores, void *ws,
                                  int64_t ws_bytes, void *stream) {
    SMZ_REQUIRE(x && p && scores && ws, "vasnet_forward: NULL pointer");
    SMZ_REQUIRE(p->wqk && p->wv && p->wo && p->w1 && p->b1 && p->w2 && p->b2 && p->ln_g && p->ln_b,
                "vasnet_forward: NULL parameter pointer");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    Plan pl;
    rc = make_plan(h_cu_seqlens, n_videos, training != 0, x_is_bf16 != 0, &pl);
    if (rc != SMZ_OK) return rc;
    // float32-accurate mode (split-bf16 operands): the lo plane of every 16-bit activation sits LO elements behind it
    const bool precise = p->wqk_lo != nullptr;
    SMZ_REQUIRE(!precise || (training && p->wv_lo && p->wo_lo && p->w1_lo),
                "vasnet_forward: the split-bf16 mode needs the training layout and all four lo weight planes");
    const int64_t LO = precise ? pl.total / 2 : 0;
    SMZ_REQUIRE(ws_bytes >= (precise ? 2 : 1) * pl.total, "vasnet_forward: work buffer too small (%lld < %lld bytes)", (long long)ws_bytes,
                (long long)((precise ? 2 : 1) * pl.total));
    SMZ_REQUIRE(training || (!drop_att && !drop_y && !drop_h), "dropout masks are only meaningful in training mode");
    const int64_t XLO = (precise && !x_is_bf16) ? LO : 0;        // bfloat16 features are exact: no lo plane
    auto wlo = [&](const void *lo, const void *hi) -> int64_t {
        return precise ? (reinterpret_cast<const char *>(lo) - reinterpret_cast<const char *>(hi)) / 2 : 0;
    };
    auto split = [](GemmEpilogue e, int64_t a_lo, int64_t b_lo, int64_t c_lo) { e.a_lo = a_lo; e.b_lo = b_lo; e.c_lo = c_lo; return e; };
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *w = reinterpret_cast<uint8_t *>(ws);
    const int n = n_videos;

    // Fast inference path (the caller passed a status word and the folded weights), see fast_chunk() below
    const bool fast = !training && !g_exact_softmax && p->status != nullptr && (x_is_bf16 ? p->wgv : p->wgv16) != nullptr &&
                      p->head_gw != nullptr && p->head_c != nullptr && p->w1g != nullptr && p->ln_c != nullptr && p->b1f != nullptr;
    std::vector<GemmProblem> probs;
    build_problems(pl, h_cu_seqlens, fast, &probs);
    GemmProblem *d_probs = reinterpret_cast<GemmProblem *>(w + pl.off_probs);
    rc = smz::upload_small(d_probs, probs.data(), probs.size() * sizeof(GemmProblem), st);
    if (rc != SMZ_OK) return rc;
    int64_t *d_dropoff = nullptr;
    if (drop_att != nullptr) {
        std::vector<int64_t> off((size_t)n);
        int64_t acc = 0;
        for (int v = 0; v < n; v++) { off[v] = acc; const int64_t T = h_cu_seqlens[v + 1] - h_cu_seqlens[v]; acc += T * T; }
        d_dropoff = reinterpret_cast<int64_t *>(w + pl.off_dropoff);
        rc = smz::upload_small(d_dropoff, off.data(), off.size() * 8, st);
        if (rc != SMZ_OK) return rc;
    }

    bf16 *qk = reinterpret_cast<bf16 *>(w + pl.off_qk);
    bf16 *vt = reinterpret_cast<bf16 *>(w + pl.off_vt);
    bf16 *o = reinterpret_cast<bf16 *>(w + pl.off_o);
    float *y = reinterpret_cast<float *>(w + pl.off_y);
    bf16 *yn = reinterpret_cast<bf16 *>(w + pl.off_yn);
    float *h = reinterpret_cast<float *>(w + pl.off_h);
    float *S = reinterpret_cast<float *>(w + pl.off_s);
    bf16 *P = reinterpret_cast<bf16 *>(w + pl.off_p);
    bf16 *alpha = (training && drop_att) ? reinterpret_cast<bf16 *>(w + pl.off_alpha) : P;
    float *stats = training ? reinterpret_cast<float *>(w + pl.off_stats) : nullptr;
    const int64_t Rs = pl.rows_cap;          // stride of the stats arrays
    bf16 *const qk0 = qk, *const vt0 = vt, *const o0 = o, *const yn0 = yn, *const P0 = P, *const alpha0 = alpha;
    float *const y0 = y, *const h0 = h, *const S0 = S, *const stats0 = stats;

    SideStreams *ss = (pl.lanes == 2 && !smz::profile_enabled()) ? side_streams(st) : nullptr;   // profiling: one stream
    const cudaStream_t caller = st;
    if (ss != nullptr) {
        SMZ_CUDA_CHECK(cudaEventRecord(ss->fork, caller));
        for (int i = 0; i < 2; i++) SMZ_CUDA_CHECK(cudaStreamWaitEvent(ss->s[i], ss->fork, 0));
    }
    int chunk_index = 0;
    for (const Chunk &c : pl.chunks) {
        const int R = c.rows;
        const int64_t Rpad = up(R, 8);
        const int64_t rb = training ? c.row0 : 0;       // row base inside the row-wise buffers
        const int lane_id = ss != nullptr ? (chunk_index & 1) : 0;
        ++chunk_index;
        if (ss != nullptr) st = ss->s[lane_id];
        const int64_t lb = (int64_t)lane_id * pl.lane_bytes;   // byte offset of this lane's buffer copy
        auto lane = [&](auto *ptr) { return reinterpret_cast<decltype(ptr)>(reinterpret_cast<uint8_t *>(ptr) + lb); };
        const int qld = training ? 2 * kFeat : 3 * kFeat;        // row length of the packed projection buffer
        qk = lane(qk0) + rb * 2 * kFeat; vt = lane(vt0) + c.vt_off; o = lane(o0) + rb * kFeat; y = lane(y0) + rb * kFeat;
        yn = lane(yn0) + rb * kFeat; h = lane(h0) + rb * kFeat; S = lane(S0) + c.lg_off; P = lane(P0) + c.lg_off;
        alpha = lane(alpha0) + c.lg_off;
        stats = training ? stats0 + rb : nullptr;
        const bf16 *xb;
        if (x_is_bf16) {
            xb = reinterpret_cast<const bf16 *>(x) + (int64_t)c.row0 * kFeat;
        } else {
            bf16 *dst = reinterpret_cast<bf16 *>(w + pl.off_xb + lb) + rb * kFeat;
            const float *src = reinterpret_cast<const float *>(x) + (int64_t)c.row0 * kFeat;
            rc = fast ? smz::launch_cvt_f16(src, dst, (int64_t)R * kFeat, p->status, SMZ_VASNET_STATUS_F16_RANGE, st)   // float16 features
                      : smz::launch_cvt_bf16(src, dst, (int64_t)R * kFeat, st, XLO);
            if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "cvt");
            xb = dst;
        }
        if (fast) {
            const void *res = x_is_bf16 ? (const void *)(reinterpret_cast<const bf16 *>(x) + (int64_t)c.row0 * kFeat)
                                        : (const void *)(reinterpret_cast<const float *>(x) + (int64_t)c.row0 * kFeat);
            rc = fast_chunk(pl, c, h_cu_seqlens, p, xb, res, x_is_bf16 != 0, qk, P, yn, reinterpret_cast<float *>(w + pl.off_sstat + lb),
                            reinterpret_cast<float *>(w + pl.off_lnstat + lb), reinterpret_cast<float *>(w + pl.off_hstat + lb),
                            d_probs, scores + c.row0, st);
            if (rc != SMZ_OK) return rc;
            continue;
        }
        // Q|K projection and V^T projection
        // training (one video, a few tiles per GEMM): V^T runs beside Q|K on the side stream, joined before alpha.V
        SideStreams *fs = (training && !smz::profile_enabled() && !smz::debug_sync_enabled()) ? side_streams(caller) : nullptr;
        cudaStream_t st_v = st;
        if (fs != nullptr) {
            SMZ_CUDA_CHECK(cudaEventRecord(fs->ev[0], st));                       // x (bf16) is ready
            SMZ_CUDA_CHECK(cudaStreamWaitEvent(fs->s[0], fs->ev[0], 0));
            st_v = fs->s[0];
        }
        smz::profile_mark(st, "gemm_qk");
        const bool wqkv_packed = reinterpret_cast<const bf16 *>(p->wv) == reinterpret_cast<const bf16 *>(p->wqk) + 2 * kFeat * kFeat;
        if (!training && wqkv_packed) {
            // inference: one [R, 3072] GEMM against [Wq; Wk; Wv] (the caller's bf16 copies are one buffer)
            rc = smz::gemm_bf16_tn(xb, R, kFeat, kFeat, p->wqk, 3 * kFeat, kFeat, kFeat, nullptr, 1, smz::gemm_tiles(R, 3 * kFeat),
                                   dense_problem(R, 3 * kFeat, kFeat, 3 * kFeat, 0), GemmEpilogue{qk, nullptr, nullptr, 1.f, 0}, st);
            if (rc != SMZ_OK) return rc;
            SMZ_DEBUG_STEP(st, "gemm_qkv");
        } else {
            rc = smz::gemm_bf16_tn(xb, R, kFeat, kFeat, p->wqk, 2 * kFeat, kFeat, kFeat, nullptr, 1, smz::gemm_tiles(R, 2 * kFeat),
                                   dense_problem(R, 2 * kFeat, kFeat, qld, 0),
                                   split(GemmEpilogue{qk, nullptr, nullptr, 1.f, 0}, XLO, wlo(p->wqk_lo, p->wqk), LO), st);
            if (rc != SMZ_OK) return rc;
            SMZ_DEBUG_STEP(st, "gemm_qk");
            smz::profile_mark(st, "gemm_vt");
            if (training)       // V^T [1024, R]
                rc = smz::gemm_bf16_tn(p->wv, kFeat, kFeat, kFeat, xb, R, kFeat, kFeat, nullptr, 1, smz::gemm_tiles(kFeat, R),
                                       dense_problem(kFeat, R, kFeat, (int)Rpad, 0),
                                       split(GemmEpilogue{vt, nullptr, nullptr, 1.f, 0}, wlo(p->wv_lo, p->wv), XLO, LO), st_v);
            else                // V into columns [2048, 3072) of the packed rows
                rc = smz::gemm_bf16_tn(xb, R, kFeat, kFeat, p->wv, kFeat, kFeat, kFeat, nullptr, 1, smz::gemm_tiles(R, kFeat),
                                       dense_problem(R, kFeat, kFeat, qld, 0), GemmEpilogue{qk + 2 * kFeat, nullptr, nullptr, 1.f, 0}, st_v);
            if (rc != SMZ_OK) return rc;
            if (fs != nullptr) SMZ_CUDA_CHECK(cudaEventRecord(fs->ev[1], st_v));
            SMZ_DEBUG_STEP(st, "gemm_vt");
        }
        // attention, one GEMM problem per video
        for (const Sub &s : c.subs) {
            const int nv = s.v1 - s.v0;
            int tiles_s = 0, tiles_pv = 0;
            for (int v = s.v0; v < s.v1; v++) {
                const int T = h_cu_seqlens[v + 1] - h_cu_seqlens[v];
                tiles_s += smz::gemm_tiles(T, T);
                tiles_pv += smz::gemm_tiles(T, kFeat);
            }
            const bool inference = !training;
            {   // fp32 logits + max-subtracted softmax (dropout in training mode)
                smz::profile_mark(st, "gemm_logits");
                const GemmEpilogue e = split(GemmEpilogue{S, nullptr, nullptr, p->scale, smz::GEMM_OUT_F32}, LO, LO, 0);
                rc = smz::gemm_bf16_tn(qk, R, qld, qld, qk, R, qld, qld, d_probs + s.v0, nv, tiles_s, GemmProblem{}, e, st);
                if (rc != SMZ_OK) return rc;
                SMZ_DEBUG_STEP(st, "gemm_logits");
                smz::profile_mark(st, "softmax");
                rc = smz::launch_softmax(d_probs + s.v0, nv, s.rows, S, alpha, P, drop_att, d_dropoff ? d_dropoff + s.v0 : nullptr,
                                         p->aperture, p->ignore_self, st, nullptr, nullptr, 0, LO);
                if (rc != SMZ_OK) return rc;
                SMZ_DEBUG_STEP(st, "softmax");
            }
            smz::profile_mark(st, "gemm_pv");
            if (fs != nullptr) SMZ_CUDA_CHECK(cudaStreamWaitEvent(st, fs->ev[1], 0));     // V^T is ready
            if (inference) {    // B = V in place: rows = frames (k), columns [2048, 3072) of the packed Q|K|V rows (MN-major)
                GemmEpilogue e{o, nullptr, nullptr, 1.f, 0};
                rc = smz::gemm_bf16(false, true, P, s.rows, s.ld, s.ld, qk, R, qld, qld, d_probs + n + s.v0, nv, tiles_pv, GemmProblem{}, e, st);
                if (rc != SMZ_OK) return rc;
            } else {
                const GemmEpilogue e = split(GemmEpilogue{o, nullptr, nullptr, 1.f, 0}, LO, LO, LO);
                rc = smz::gemm_bf16_tn(P, s.rows, s.ld, s.ld, vt, kFeat, R, Rpad, d_probs + n + s.v0, nv, tiles_pv, GemmProblem{}, e, st);
                if (rc != SMZ_OK) return rc;
            }
        SMZ_DEBUG_STEP(st, "gemm_pv");
        }
        // output projection + residual, LayerNorm, k1 + ReLU, head
        const void *res = x_is_bf16 ? (const void *)(reinterpret_cast<const bf16 *>(x) + (int64_t)c.row0 * kFeat)
                                    : (const void *)(reinterpret_cast<const float *>(x) + (int64_t)c.row0 * kFeat);
        const bool fused_head = !training && p->head_gw != nullptr && p->head_c != nullptr;
        smz::profile_mark(st, "gemm_out");
        rc = smz::gemm_bf16_tn(o, R, kFeat, kFeat, p->wo, kFeat, kFeat, kFeat, nullptr, 1, smz::gemm_tiles(R, kFeat),
                               dense_problem(R, kFeat, kFeat, kFeat, kFeat),
                               split(GemmEpilogue{y, nullptr, res, 1.f, smz::GEMM_OUT_F32 | (x_is_bf16 ? 0 : smz::GEMM_RES_F32)},
                                     LO, wlo(p->wo_lo, p->wo), 0), st);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "gemm_out");
        smz::profile_mark(st, "layernorm");
        rc = smz::launch_layernorm(y, drop_y ? drop_y + (int64_t)c.row0 * kFeat : nullptr, p->ln_g, p->ln_b, p->eps, R, yn,
                                   stats, stats ? stats + Rs : nullptr, st, LO);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "layernorm");
        smz::profile_mark(st, "gemm_k1");
        if (fused_head) {
            // inference: H never reaches memory — the k1 epilogue reduces relu(h) to the three row sums the second
            // LayerNorm + k2 dot need (GEMM_ROWSTATS), a one-thread-per-frame kernel finishes the score
            float *hstat = reinterpret_cast<float *>(w + pl.off_hstat + lb);
            GemmEpilogue e{h, p->b1, nullptr, 1.f, smz::GEMM_RELU | smz::GEMM_OUT_F32 | smz::GEMM_ROWSTATS | smz::GEMM_NO_STORE};
            e.stat_w = p->head_gw; e.stat_out = hstat;
            rc = smz::gemm_bf16_tn(yn, R, kFeat, kFeat, p->w1, kFeat, kFeat, kFeat, nullptr, 1, smz::gemm_tiles(R, kFeat),
                                   dense_problem(R, kFeat, kFeat, kFeat, 0), e, st);
            if (rc != SMZ_OK) return rc;
            SMZ_DEBUG_STEP(st, "gemm_k1");
            smz::profile_mark(st, "head");
            rc = smz::launch_head_from_stats(hstat, 2 * kFeat / smz::GEMM_BN, p->head_c, p->eps, R, scores + c.row0, st);
            if (rc != SMZ_OK) return rc;
        } else {
        rc = smz::gemm_bf16_tn(yn, R, kFeat, kFeat, p->w1, kFeat, kFeat, kFeat, nullptr, 1, smz::gemm_tiles(R, kFeat),
                               dense_problem(R, kFeat, kFeat, kFeat, 0),
                               split(GemmEpilogue{h, p->b1, nullptr, 1.f, smz::GEMM_RELU | smz::GEMM_OUT_F32}, LO, wlo(p->w1_lo, p->w1), 0), st);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "gemm_k1");
        smz::profile_mark(st, "head");
        rc = smz::launch_head(h, drop_h ? drop_h + (int64_t)c.row0 * kFeat : nullptr, p->ln_g, p->ln_b, p->eps, p->w2, p->b2,
                              R, scores + c.row0, stats ? stats + 2 * Rs : nullptr, stats ? stats + 3 * Rs : nullptr, st);
        if (rc != SMZ_OK) return rc;
        }
        SMZ_DEBUG_STEP(st, "head");
        smz::profile_mark(st, "");
    }
    if (ss != nullptr) {
        for (int i = 0; i < 2; i++) {
            SMZ_CUDA_CHECK(cudaEventRecord(ss->join[i], ss->s[i]));
            SMZ_CUDA_CHECK(cudaStreamWaitEvent(caller, ss->join[i], 0));
        }
    }
    return SMZ_OK;
}

// ---------------------------------------------------------------------------------------------------
// backward: the autograd graph behind vasnet.py:209-211 (loss.backward()) for the scorer, video by video,
// on the activations smz_vasnet_forward(training=1) left in the work buffer.  Every contraction runs on
// the same tcgen05 GEMM; MN-major operand forms read the row-major activations / weights directly:
//   dYn = dh . W1            dW1  += dh^T . Yn
//   dO  = dY . Wo            dWo  += dY^T . O
//   dP  = dO . V^T           dV^T  = dO^T . P
//   dQ  = s dS . K           dK    = s dS^T . Q          (s = scale)
//   dWqk += [dQ|dK]^T . xb   dWv  += dV^T . xb           dx = dY + [dQ|dK] . Wqk + dV . Wv   (optional)
// Gradients are ACCUMULATED (+=) into the caller's float32 buffers.
namespace {

int gemm1(bool a_mn, bool b_mn, const void *A, int64_t a_rows, int64_t a_cols, int64_t lda, const void *B,
          int64_t b_rows, int64_t b_cols, int64_t ldb, GemmProblem g, const GemmEpilogue &e, cudaStream_t st) {
    g.tiles_n = (g.N + smz::GEMM_BN - 1) / smz::GEMM_BN;
    return smz::gemm_bf16(a_mn, b_mn, A, a_rows, a_cols, lda, B, b_rows, b_cols, ldb, nullptr, 1, smz::gemm_tiles(g.M, g.N), g, e, st);
}

GemmProblem prob(int M, int N, int K, int ldc, int64_t c_off = 0, int ldr = 0, int64_t r_off = 0) {
    GemmProblem g = {};
    g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.c_off = c_off; g.ldr = ldr; g.r_off = r_off;
    return g;
}

}  // namespace

extern "C" int smz_vasnet_backward(const void *x, int x_is_bf16, const int32_t *h_cu_seqlens, int n_videos,
                                   const smz_vasnet_params *p, const uint8_t *drop_att, const uint8_t *drop_y,
                                   const uint8_t *drop_h, const float *scores, const float *dscores,
                                   const smz_vasnet_grads *gr, void *ws, int64_t ws_bytes, void *stream) {
    SMZ_REQUIRE(x && p && scores && dscores && gr && ws, "vasnet_backward: NULL pointer");
    SMZ_REQUIRE(gr->wqk && gr->wv && gr->wo && gr->w1 && gr->b1 && gr->w2 && gr->b2 && gr->ln_g && gr->ln_b,
                "vasnet_backward: NULL gradient pointer");
    int rc = smz_device_check();
    if (rc != SMZ_OK) return rc;
    Plan pl;
    rc = make_plan(h_cu_seqlens, n_videos, true, x_is_bf16 != 0, &pl);
    if (rc != SMZ_OK) return rc;
    const bool precise = p->wqk_lo != nullptr;                   // float32-accurate mode, see smz_vasnet_forward
    SMZ_REQUIRE(!precise || (p->wv_lo && p->wo_lo && p->w1_lo), "vasnet_backward: the split-bf16 mode needs all four lo weight planes");
    const int64_t LO = precise ? pl.total / 2 : 0;
    const int64_t XLO = (precise && !x_is_bf16) ? LO : 0;
    SMZ_REQUIRE(ws_bytes >= (precise ? 2 : 1) * pl.total, "vasnet_backward: work buffer too small (%lld < %lld bytes)", (long long)ws_bytes,
                (long long)((precise ? 2 : 1) * pl.total));
    auto wlo = [&](const void *lo, const void *hi) -> int64_t {
        return precise ? (reinterpret_cast<const char *>(lo) - reinterpret_cast<const char *>(hi)) / 2 : 0;
    };
    auto split = [](GemmEpilogue e, int64_t a_lo, int64_t b_lo, int64_t c_lo) { e.a_lo = a_lo; e.b_lo = b_lo; e.c_lo = c_lo; return e; };
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *w = reinterpret_cast<uint8_t *>(ws);
    const int64_t Rs = pl.rows_cap;
    const int F32ACC = smz::GEMM_OUT_F32 | smz::GEMM_RES_F32;   // C = A.B^T + C (float32)
    int64_t att_off = 0;                                         // offset of the video's attention keep-mask
    // A batch-1 backward pass is 12 GEMMs of a few tiles each: the weight-gradient GEMMs (and dV^T, dK) do not sit on the
    // dX chain, so they run on a side stream beside it — fork / join with events, which a CUDA-graph capture records as
    // parallel branches.  main: head, dYn, LN, dO, dP, softmax, dQ | side: dW1, dWo, dV^T, dWv, dK | join | dWqk, dx.
    SideStreams *ss = (!smz::profile_enabled() && !smz::debug_sync_enabled()) ? side_streams(st) : nullptr;
    cudaStream_t sd = ss != nullptr ? ss->s[0] : st;
    auto fork_to_side = [&](int e) -> int {                      // side stream continues after what main has queued so far
        if (ss == nullptr) return SMZ_OK;
        SMZ_CUDA_CHECK(cudaEventRecord(ss->ev[e], st));
        SMZ_CUDA_CHECK(cudaStreamWaitEvent(sd, ss->ev[e], 0));
        return SMZ_OK;
    };
    rc = fork_to_side(0);                                        // the caller's zeroed gradient buffers
    if (rc != SMZ_OK) return rc;

    for (const Chunk &c : pl.chunks) {
        const int T = c.rows;
        const int64_t rb = c.row0, Tpad = up(T, 8);
        const int ld = c.subs[0].ld;
        const bf16 *xb = x_is_bf16 ? reinterpret_cast<const bf16 *>(x) + rb * kFeat
                                   : reinterpret_cast<const bf16 *>(w + pl.off_xb) + rb * kFeat;
        const bf16 *qk = reinterpret_cast<const bf16 *>(w + pl.off_qk) + rb * 2 * kFeat;
        const bf16 *vt = reinterpret_cast<const bf16 *>(w + pl.off_vt) + c.vt_off;
        const bf16 *o = reinterpret_cast<const bf16 *>(w + pl.off_o) + rb * kFeat;
        const float *y = reinterpret_cast<const float *>(w + pl.off_y) + rb * kFeat;
        const bf16 *yn = reinterpret_cast<const bf16 *>(w + pl.off_yn) + rb * kFeat;
        const float *h = reinterpret_cast<const float *>(w + pl.off_h) + rb * kFeat;
        const bf16 *P = reinterpret_cast<const bf16 *>(w + pl.off_p) + c.lg_off;
        const bf16 *alpha = drop_att ? reinterpret_cast<const bf16 *>(w + pl.off_alpha) + c.lg_off : P;
        const float *stats = reinterpret_cast<const float *>(w + pl.off_stats) + rb;
        bf16 *dh = reinterpret_cast<bf16 *>(w + pl.off_dh) + rb * kFeat;
        float *dyn = reinterpret_cast<float *>(w + pl.off_dyn) + rb * kFeat;
        bf16 *dy = reinterpret_cast<bf16 *>(w + pl.off_dy) + rb * kFeat;
        float *dyf = gr->dx ? reinterpret_cast<float *>(w + pl.off_dyf) + rb * kFeat : nullptr;
        bf16 *dO = reinterpret_cast<bf16 *>(w + pl.off_do) + rb * kFeat;
        bf16 *dqk = reinterpret_cast<bf16 *>(w + pl.off_dqk) + rb * 2 * kFeat;
        bf16 *dvt = reinterpret_cast<bf16 *>(w + pl.off_dvt) + c.vt_off;
        float *dP = reinterpret_cast<float *>(w + pl.off_dp) + c.lg_off;
        bf16 *dS = reinterpret_cast<bf16 *>(w + pl.off_ds) + c.lg_off;

        // regressor head, second LayerNorm, dropout, ReLU  ->  dh (k1 pre-activation gradient)
        rc = smz::launch_head_bwd(h, drop_h ? drop_h + rb * kFeat : nullptr, p->ln_g, p->ln_b, p->w2, stats + 2 * Rs,
                                  stats + 3 * Rs, scores + rb, dscores + rb, T, dh, gr->w2, gr->b2, gr->ln_g, gr->ln_b,
                                  gr->b1, st, LO);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "head_bwd");
        rc = fork_to_side(1);                                    // dh is ready
        if (rc != SMZ_OK) return rc;
        // k1: dYn = dh . W1 ; dW1 += dh^T . Yn
        rc = gemm1(false, true, dh, T, kFeat, kFeat, p->w1, kFeat, kFeat, kFeat, prob(T, kFeat, kFeat, kFeat),
                   split(GemmEpilogue{dyn, nullptr, nullptr, 1.f, smz::GEMM_OUT_F32}, LO, wlo(p->w1_lo, p->w1), 0), st);
        if (rc != SMZ_OK) return rc;
        rc = gemm1(true, true, dh, T, kFeat, kFeat, yn, T, kFeat, kFeat, prob(kFeat, kFeat, T, kFeat, 0, kFeat),
                   split(GemmEpilogue{gr->w1, nullptr, gr->w1, 1.f, F32ACC}, LO, LO, 0), sd);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "k1_bwd");
        // first LayerNorm + dropout -> dY (also the gradient of the residual branch)
        rc = smz::launch_layernorm_bwd(dyn, y, drop_y ? drop_y + rb * kFeat : nullptr, p->ln_g, stats, stats + Rs, T, dy, dyf,
                                       gr->ln_g, gr->ln_b, st, LO);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "ln_bwd");
        rc = fork_to_side(2);                                    // dY is ready
        if (rc != SMZ_OK) return rc;
        // output projection: dO = dY . Wo ; dWo += dY^T . O
        rc = gemm1(false, true, dy, T, kFeat, kFeat, p->wo, kFeat, kFeat, kFeat, prob(T, kFeat, kFeat, kFeat),
                   split(GemmEpilogue{dO, nullptr, nullptr, 1.f, 0}, LO, wlo(p->wo_lo, p->wo), LO), st);
        if (rc != SMZ_OK) return rc;
        rc = gemm1(true, true, dy, T, kFeat, kFeat, o, T, kFeat, kFeat, prob(kFeat, kFeat, T, kFeat, 0, kFeat),
                   split(GemmEpilogue{gr->wo, nullptr, gr->wo, 1.f, F32ACC}, LO, LO, 0), sd);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "out_bwd");
        rc = fork_to_side(3);                                    // dO is ready
        if (rc != SMZ_OK) return rc;
        // attention: dP = dO . V^T (B = V^T stored [d][j]: MN-major) ; dV^T = dO^T . P ; dWv += dV^T . xb
        rc = gemm1(false, true, dO, T, kFeat, kFeat, vt, kFeat, T, Tpad, prob(T, T, kFeat, ld),
                   split(GemmEpilogue{dP, nullptr, nullptr, 1.f, smz::GEMM_OUT_F32}, LO, LO, 0), st);
        if (rc != SMZ_OK) return rc;
        rc = gemm1(true, true, dO, T, kFeat, kFeat, P, T, T, ld, prob(kFeat, T, T, (int)Tpad),
                   split(GemmEpilogue{dvt, nullptr, nullptr, 1.f, 0}, LO, LO, LO), sd);
        if (rc != SMZ_OK) return rc;
        rc = gemm1(false, true, dvt, kFeat, T, Tpad, xb, T, kFeat, kFeat, prob(kFeat, kFeat, T, kFeat, 0, kFeat),
                   split(GemmEpilogue{gr->wv, nullptr, gr->wv, 1.f, F32ACC}, LO, XLO, 0), sd);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "att_bwd1");
        rc = smz::launch_softmax_bwd(dP, alpha, drop_att ? drop_att + att_off : nullptr, T, ld, dS, st, LO);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "softmax_bwd");
        rc = fork_to_side(4);                                    // dS is ready
        if (rc != SMZ_OK) return rc;
        // dQ = scale dS . K (B = K half of QK, stored [j][d]: MN-major at column 1024) ; dK = scale dS^T . Q
        {
            GemmProblem g = prob(T, kFeat, T, 2 * kFeat, 0);
            g.b_col0 = kFeat;
            rc = gemm1(false, true, dS, T, T, ld, qk, T, 2 * kFeat, 2 * kFeat, g, split(GemmEpilogue{dqk, nullptr, nullptr, p->scale, 0}, LO, LO, LO), st);
            if (rc != SMZ_OK) return rc;
            g = prob(T, kFeat, T, 2 * kFeat, kFeat);
            rc = gemm1(true, true, dS, T, T, ld, qk, T, 2 * kFeat, 2 * kFeat, g, split(GemmEpilogue{dqk, nullptr, nullptr, p->scale, 0}, LO, LO, LO), sd);
            if (rc != SMZ_OK) return rc;
        }
        SMZ_DEBUG_STEP(st, "att_bwd2");
        if (ss != nullptr) {                                     // join: dK, dV^T and the side weight gradients are done
            SMZ_CUDA_CHECK(cudaEventRecord(ss->ev[5], sd));
            SMZ_CUDA_CHECK(cudaStreamWaitEvent(st, ss->ev[5], 0));
        }
        // projections: dWqk += [dQ|dK]^T . xb
        rc = gemm1(true, true, dqk, T, 2 * kFeat, 2 * kFeat, xb, T, kFeat, kFeat, prob(2 * kFeat, kFeat, T, kFeat, 0, kFeat),
                   split(GemmEpilogue{gr->wqk, nullptr, gr->wqk, 1.f, F32ACC}, LO, XLO, 0), st);
        if (rc != SMZ_OK) return rc;
        SMZ_DEBUG_STEP(st, "proj_bwd");
        if (gr->dx != nullptr) {
            float *dx = gr->dx + rb * kFeat;
            // dx = dY + [dQ|dK] . Wqk   (B = Wqk stored [c][i]: MN-major), then += dV . Wv (A = dV^T: MN-major)
            rc = gemm1(false, true, dqk, T, 2 * kFeat, 2 * kFeat, p->wqk, 2 * kFeat, kFeat, kFeat,
                       prob(T, kFeat, 2 * kFeat, kFeat, 0, kFeat), split(GemmEpilogue{dx, nullptr, dyf, 1.f, F32ACC}, LO, wlo(p->wqk_lo, p->wqk), 0), st);
            if (rc != SMZ_OK) return rc;
            rc = gemm1(true, true, dvt, kFeat, T, Tpad, p->wv, kFeat, kFeat, kFeat, prob(T, kFeat, kFeat, kFeat, 0, kFeat),
                       split(GemmEpilogue{dx, nullptr, dx, 1.f, F32ACC}, LO, wlo(p->wv_lo, p->wv), 0), st);
            if (rc != SMZ_OK) return rc;
            SMZ_DEBUG_STEP(st, "dx");
        }
        att_off += (int64_t)T * T;
    }
    return SMZ_OK;
}
