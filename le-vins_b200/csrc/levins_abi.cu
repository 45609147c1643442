// levins_abi.cu -- host side of liblevins_b200.so: context, device memory, launch sequencing, the C ABI.
// No torch, no third-party libraries; CUDA runtime only.  There is no CPU fallback: every compute entry point
// launches the sm_100a kernels of kernels.cuh and reports LEVINS_ERR_CUDA when that is impossible.
#include "levins/abi.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cfloat>
#include <cstring>
#include <deque>
#include <initializer_list>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace lv;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        size_t alloc = need + need / 4 + 256;
        cudaError_t e = cudaMalloc(&p, alloc);
        if (e == cudaSuccess) bytes = alloc;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

} // namespace

struct levins_ctx {
    int device = 0;
    levins_config cfg{};
    Consts consts{};
    Extrinsics ext{};
    bool have_ext = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    std::mutex mu; // append_scan may come from another thread
    uint64_t launches = 0;
    int num_sms = 148;
    bool deskew_attr_set = false;
    DevBuf cull_stat;              // {tried, fired} chunk counts of the launches since the last read-back
    bool cull_off = false;         // the cull never fired lately: run the instantiation without it
    unsigned cull_off_launches = 0;
    unsigned kf_calls = 0;
    cudaStream_t side = nullptr; // k_count_cells runs here, overlapped with the latency-bound k_associate
    cudaEvent_t ev_binned = nullptr, ev_counted = nullptr;
    cudaEvent_t ev_append = nullptr; // recorded after the host->device copies of the latest levins_append_scan / levins_append_raw

    // ---- streaming state ----
    DevBuf ring;      // PointXYZIT x cfg.max_points
    DevBuf ring_state; // RingState
    uint64_t ring_tail = 0;     // physical index one past the newest point (levins_append_scan's copy target)
    bool ring_tail_known = true; // false after levins_append_raw until the ring state is read back
    uint64_t ring_count_ub = 0; // host's upper bound of the live point count
    DevBuf s_ins, s_table, s_kf, s_queries;
    DevBuf chunks; // ChunkDesc of the current k_deskew_bin launch (streaming and batch)
    DevBuf raw_stage, raw_blocks, raw_misc; // levins_append_raw: records, per-block counts, {min, max, result}
    DevBuf s_type, s_xyzd, s_flags, s_knn_idx, s_knn_d2, s_nn, s_nn_meta;
    DevBuf s_pixels, s_undist, s_accepted, s_pw; // feature side on the device (levins_associate_pixels)
    DevBuf scratch; // inputs / outputs of the small batched entry points (plane geometry, depth factors)
    DevBuf s_arena_in, s_arena_out; // fused keyframe: everything sent / returned in one copy each (keyframe_body)
    void* pin_io = nullptr;       // their page-locked mirrors (first half in, second half out)
    size_t pin_io_bytes = 0;
    DevBuf s_cloud;    // resident deskewed cloud (PointXYZI)
    DevBuf s_in_cloud; // explicit cloud uploaded by build_range_image
    DevBuf s_zbuf, s_cells, s_c2c, s_unit, s_winner;
    unsigned s_gen = 255; // z-buffer generation of the streaming slot (255 = never initialised)
    unsigned s_img_gen = 0; // generation whose winners are in s_zbuf / s_cells
    size_t s_cloud_n = 0;     // points in the resident cloud
    bool s_cloud_valid = false;
    bool s_image_valid = false;
    bool s_image_from_raw = false; // cells were built from the ring (fused) rather than from a cloud
    // pinned staging area of the stage-wise entry points: [0,256) KfState | [256,512) RingState | [512,1024) misc | INS window.
    // Small host arrays go through it so that their copies are truly asynchronous (copies from / to pageable memory block the
    // calling thread one by one).  The fused keyframe has its own arenas (pin_io).
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    size_t pin_ins() const { return 1024; }
    size_t pin_total() const { return pin_ins() + ((cfg.max_ins * sizeof(InsSample) + 255) & ~size_t(255)); }

    // ---- batch state ----
    DevBuf b_pts, b_ins, b_table, b_queries, b_kfs;
    DevBuf b_results; // xyz_depth (16 B / query) immediately followed by type (1 B / query): one contiguous block, so that an
                      // exchange step can move both with a single collective
    size_t b_results_nq = 0;
    DevBuf b_flags, b_nn, b_nn_meta;
    DevBuf b_zbuf, b_cells;
    std::vector<levins_keyframe_desc> b_descs;
    std::vector<KfState> b_kfs_host;
    // levins_batch_attach_device: the pools are the caller's device memory (null: the library's own b_pts / b_ins / b_queries)
    const void *b_pts_view = nullptr, *b_ins_view = nullptr, *b_queries_view = nullptr;
    const PointXYZIT* batch_pts() const { return static_cast<const PointXYZIT*>(b_pts_view ? b_pts_view : b_pts.p); }
    const InsSample* batch_ins() const { return static_cast<const InsSample*>(b_ins_view ? b_ins_view : b_ins.p); }
    const float* batch_queries() const { return static_cast<const float*>(b_queries_view ? b_queries_view : b_queries.p); }
    DevBuf b_packed, b_scans; // levins_batch_upload_packed: the records as uploaded, the scan stamps
    size_t b_nq = 0, b_npts = 0, b_nins = 0;
    int b_wave = 256;
    int b_wave_alloc = 0;
    unsigned b_gen = 255;
    bool b_valid = false;
    // ---- exchange step (exchange.cuh): gathered results of all ranks ----
    DevBuf g_block;            // [PEER_HDR | halves x ranks x slot]: NCCL mode one half, peer mode two (step parity)
    size_t g_slot = 0, g_nq = 0; // bytes / queries per rank slot
    int g_ranks = 0, g_half = 0;
    PeerOut peer{};
    PeerFlags pflags{};
    void* peer_mapped[MAX_PEERS] = {};
    unsigned long long peer_step = 0;
    bool peer_on = false;
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    struct EvSpan {
        int kind;
        size_t e0, e1;
    };
    std::vector<EvSpan> ev_spans;
    float prof_ms[4] = {0, 0, 0, 0};
    uint32_t prof_launches[4] = {0, 0, 0, 0};
};

namespace {

int fail(levins_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}
int cuda_fail(levins_ctx* c, cudaError_t e, const char* what) {
    return fail(c, LEVINS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                                    \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call);  \
    } while (0)

// tracking.h:119-135 evaluated with the reference's own double expressions, on the host's libm like the reference
Consts make_consts(const levins_config& cfg) {
    Consts c{};
    const double ANGLE = 0.3 * M_PI / 180.0;
    const double FOV = 90.0 * M_PI / 180.0;
    c.angle = ANGLE;
    c.inv_angle = 1.0 / ANGLE;
    c.half_offset = M_PI * 0.5 - FOV * 0.5;
    c.tan_value = tan(FOV * 0.5);
    const double thr = sin(ANGLE) * cfg.nearest_distance_counts;
    c.thr2 = pow(thr, 2);
    c.nearest = cfg.nearest_depth;
    c.farthest = cfg.farthest_depth;
    c.plane_residual = cfg.plane_residual_threshold;
    c.range_spread = cfg.range_spread_threshold;
    c.stack_interval = cfg.stack_interval;
    c.libm_mode = cfg.libm_mode;
    // Range gate restated on the squared norm: sqrtf is monotone and correctly rounded on host and device alike, so
    // "(double) sqrtf(n2) > farthest" flips at one float; find it by bisection over the (ordered) positive float patterns.
    auto as_f = [](uint32_t b) { float f; memcpy(&f, &b, 4); return f; };
    {
        uint32_t lo = 0, hi = 0x7f7fffffu; // largest n2 that is NOT gated as "too far"
        if ((double) sqrtf(as_f(hi)) > c.farthest) {
            while (hi - lo > 1) {
                const uint32_t mid = lo + (hi - lo) / 2;
                if ((double) sqrtf(as_f(mid)) > c.farthest) hi = mid; else lo = mid;
            }
            c.n2_far = (double) sqrtf(as_f(lo)) > c.farthest ? -1.0f : as_f(lo);
        } else {
            c.n2_far = as_f(hi);
        }
    }
    {
        uint32_t lo = 0, hi = 0x7f7fffffu; // smallest n2 that is NOT gated as "too near"
        if ((double) sqrtf(as_f(lo)) < c.nearest) {
            while (hi - lo > 1) {
                const uint32_t mid = lo + (hi - lo) / 2;
                if ((double) sqrtf(as_f(mid)) < c.nearest) lo = mid; else hi = mid;
            }
            c.n2_near = as_f(hi);
        } else {
            c.n2_near = 0.0f;
        }
    }
    c.inv_angle_f = (float) (1.0 / ANGLE);
    c.cell_guard = 4e-4f; // >= 4x the worst-case error of the polynomial cell path
    return c;
}

// Start a new z-buffer generation for `n` cells (see key_make): a real clear only every 254 uses.
int next_gen(levins_ctx* ctx, unsigned long long* z, size_t n, unsigned* gen) {
    if (*gen <= 1 || *gen > 254) {
        CU(cudaMemsetAsync(z, 0xFF, n * sizeof(unsigned long long), ctx->stream));
        *gen = 254;
    } else {
        *gen -= 1;
    }
    return LEVINS_OK;
}

int ensure_stream_buffers(levins_ctx* ctx, size_t n_ins, size_t n_feat) {
    CU(ctx->s_ins.ensure(std::max<size_t>(n_ins, 1) * sizeof(InsSample)));
    CU(ctx->s_table.ensure((std::max<size_t>(n_ins, 1) + 1) * sizeof(Interval))); // + the fallback copy
    CU(ctx->s_kf.ensure(sizeof(KfState)));
    CU(ctx->s_queries.ensure(std::max<size_t>(n_feat, 1) * 3 * sizeof(float)));
    CU(ctx->s_type.ensure(std::max<size_t>(n_feat, 1)));
    CU(ctx->s_xyzd.ensure(std::max<size_t>(n_feat, 1) * 16));
    CU(ctx->s_flags.ensure(std::max<size_t>(n_feat, 1) * 4));
    CU(ctx->s_knn_idx.ensure(std::max<size_t>(n_feat, 1) * 20));
    CU(ctx->s_knn_d2.ensure(std::max<size_t>(n_feat, 1) * 20));
    CU(ctx->s_nn.ensure(std::max<size_t>(n_feat, 1) * 80));
    CU(ctx->s_nn_meta.ensure(std::max<size_t>(n_feat, 1) * 8));
    CU(ctx->s_pixels.ensure(std::max<size_t>(n_feat, 1) * 8));
    CU(ctx->s_undist.ensure(std::max<size_t>(n_feat, 1) * 8));
    CU(ctx->s_accepted.ensure(std::max<size_t>(n_feat, 1)));
    CU(ctx->s_pw.ensure(std::max<size_t>(n_feat, 1) * 24));
    CU(ctx->s_cells.ensure((size_t) CELLS * sizeof(float4)));
    CU(ctx->s_c2c.ensure((size_t) CELLS * 4));
    CU(ctx->s_unit.ensure((size_t) CELLS * sizeof(PointXYZI)));
    CU(ctx->s_winner.ensure((size_t) CELLS * 4));
    if (!ctx->s_zbuf.p) {
        CU(ctx->s_zbuf.ensure((size_t) CELLS * 8));
        ctx->s_gen = 255;
    }
    return LEVINS_OK;
}

int ensure_ring(levins_ctx* ctx) {
    if (ctx->ring.p) return LEVINS_OK;
    CU(ctx->ring.ensure((size_t) ctx->cfg.max_points * sizeof(PointXYZIT)));
    ctx->ring.bytes = (size_t) ctx->cfg.max_points * sizeof(PointXYZIT);
    CU(ctx->ring_state.ensure(sizeof(RingState)));
    CU(cudaMemsetAsync(ctx->ring_state.p, 0, sizeof(RingState), ctx->stream));
    ctx->ring_tail = 0;
    ctx->ring_tail_known = true;
    ctx->ring_count_ub = 0;
    return LEVINS_OK;
}

// Chunk descriptors for k_deskew_bin (written by k_prepare): size the buffer for `n_kf` keyframes of at most `max_pts` points
int size_chunks(levins_ctx* ctx, DeskewArgs& a, uint64_t max_pts, unsigned n_kf) {
    a.n_kf = n_kf;
    a.chunks_per_kf = (unsigned) ((max_pts + CHUNK_PTS - 1) / CHUNK_PTS);
    const uint64_t total = (uint64_t) a.n_kf * a.chunks_per_kf;
    if (total >= (1ull << 30)) return fail(ctx, LEVINS_ERR_CAPACITY, "more than 2^30 chunks in one launch: lower the wave size");
    CU(ctx->chunks.ensure(std::max<uint64_t>(total, 1) * sizeof(ChunkDesc)));
    a.chunks = ctx->chunks.as<ChunkDesc>();
    return LEVINS_OK;
}
// k_prepare's grid: one row of blocks per keyframe, enough slices that a block writes at most LV_PREP_CHUNKS descriptors, and at
// least a few slices when there are few keyframes (the interval table is split over them as well)
dim3 prepare_grid(const levins_ctx* ctx, unsigned n_kf, unsigned chunks_per_kf) {
    unsigned slices = std::max(1u, (chunks_per_kf + (unsigned) LV_PREP_CHUNKS - 1u) / (unsigned) LV_PREP_CHUNKS);
    const unsigned room = (unsigned) (2 * ctx->num_sms) / std::max(1u, n_kf); // blocks per keyframe that still run in one wave
    slices = std::max(slices, std::min(4u, room));
    return dim3(n_kf, std::min(slices, 65535u));
}

// k_deskew_bin is persistent: one 16-warp CTA per SM (221 KB of TMA staging), every warp walking its own chunk list.
// k_prepare must have filled a.chunks / a.chunks_per_kf.
int launch_deskew(levins_ctx* ctx, DeskewArgs& a, bool write_cloud) {
    const unsigned total = a.n_kf * a.chunks_per_kf;
    if (total == 0) return LEVINS_OK;
    using Kernel = void (*)(DeskewArgs);
    static const Kernel kernels[2][2][2] = {{{k_deskew_bin<0, false, false>, k_deskew_bin<0, false, true>}, {k_deskew_bin<0, true, false>, k_deskew_bin<0, true, false>}},
                                            {{k_deskew_bin<1, false, false>, k_deskew_bin<1, false, true>}, {k_deskew_bin<1, true, false>, k_deskew_bin<1, true, false>}}};
    if (!ctx->deskew_attr_set) {
        for (int i = 0; i < 8; i++)
            CU(cudaFuncSetAttribute(kernels[i >> 2][(i >> 1) & 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, DESKEW_SMEM));
        CU(ctx->cull_stat.ensure(16));
        CU(cudaMemsetAsync(ctx->cull_stat.p, 0, 16, ctx->stream));
        ctx->deskew_attr_set = true;
    }
    // The certified frustum cull pays for 360-degree scanners and costs ~5 % where every point lies inside the image cone.
    // The kernel reports how often it fired; when it (almost) never does, the instantiation without it runs, with a probe
    // every 64th launch.  Both give identical results.
    bool cull = !write_cloud;
    if (cull && ctx->cull_off) {
        if (++ctx->cull_off_launches >= 64) {
            ctx->cull_off = false; // probe again
            ctx->cull_off_launches = 0;
        } else {
            cull = false;
        }
    }
    a.cull_stat = cull ? ctx->cull_stat.as<unsigned>() : nullptr;
    const unsigned grid = std::min<unsigned>((total + DESKEW_WARPS - 1) / DESKEW_WARPS, (unsigned) LV_MINBLK * (unsigned) ctx->num_sms);
    kernels[ctx->consts.libm_mode ? 1 : 0][write_cloud ? 1 : 0][cull ? 1 : 0]<<<grid, DESKEW_BLOCK, DESKEW_SMEM, ctx->stream>>>(a);
    ctx->launches++;
    CU(cudaGetLastError());
    return LEVINS_OK;
}

// Call right after a synchronisation that follows deskew launches: fetch and reset the cull statistics (8 bytes) and decide
// which instantiation the next launches use.
int update_cull_policy(levins_ctx* ctx) {
    if (!ctx->cull_stat.p || ctx->cull_off) return LEVINS_OK;
    unsigned st[2] = {0, 0};
    CU(cudaMemcpyAsync(st, ctx->cull_stat.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemsetAsync(ctx->cull_stat.p, 0, 8, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (st[0] >= 64 && (double) st[1] < 0.02 * (double) st[0]) {
        ctx->cull_off = true;
        ctx->cull_off_launches = 0;
    }
    return LEVINS_OK;
}

// read the ring state (synchronises)
int read_ring_state(levins_ctx* ctx, RingState* out) {
    CU(cudaMemcpyAsync(ctx->pinned, ctx->ring_state.p, sizeof(RingState), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->pinned, sizeof(RingState));
    ctx->ring_count_ub = out->count;
    return LEVINS_OK;
}

// Where the per-keyframe inputs and outputs of the streaming path live on the device.  The stage-wise entry points use the
// context's individual buffers; the fused keyframe packs everything into one input and one output arena so that a keyframe
// costs one host->device and one device->host copy.
struct StreamPtrs {
    KfState* kf;
    InsSample* ins;
    float* queries;
    float* pixels;
    int8_t* type;
    float4* xyzd;
    unsigned char* accepted;
    float* undist;
    double* pw;
    KfState* kf_out = nullptr; // arena mode: k_plane leaves the final states next to the results
    RingState* ring_out = nullptr;
};
StreamPtrs default_ptrs(levins_ctx* ctx) {
    StreamPtrs sp{};
    sp.kf = ctx->s_kf.as<KfState>();
    sp.ins = ctx->s_ins.as<InsSample>();
    sp.queries = ctx->s_queries.as<float>();
    sp.pixels = ctx->s_pixels.as<float>();
    sp.type = ctx->s_type.as<int8_t>();
    sp.xyzd = ctx->s_xyzd.as<float4>();
    sp.accepted = ctx->s_accepted.as<unsigned char>();
    sp.undist = ctx->s_undist.as<float>();
    sp.pw = ctx->s_pw.as<double>();
    return sp;
}

// stage 1 on the ring: k_prepare (+ pops, + chunk descriptors) then k_deskew_bin.  Leaves KfState in sp.kf.
// uploaded: the caller has already put KfState and the INS window on the device (arena mode).
int run_stage1(levins_ctx* ctx, const StreamPtrs& sp, double stamp, const levins_ins_sample* ins, size_t n_ins, bool write_cloud,
               size_t n_queries = 0, bool uploaded = false) {
    if (!ctx->have_ext) return fail(ctx, LEVINS_ERR_STATE, "levins_set_extrinsics has not been called");
    if (!ins || n_ins == 0) return fail(ctx, LEVINS_ERR_ARG, "empty INS window");
    if (n_ins > ctx->cfg.max_ins) return fail(ctx, LEVINS_ERR_CAPACITY, "INS window larger than config.max_ins");
    if (!ctx->ring.p || ctx->ring_count_ub == 0) return LEVINS_EMPTY; // le_vins.cc:378: lidar_point_buffer_.empty()
    // everything that can refuse the call comes before k_prepare, which pops the ring (shared state)
    const size_t n_ub = ctx->ring_count_ub;
    if (n_ub >= MAX_POINTS_PER_KEYFRAME) return fail(ctx, LEVINS_ERR_CAPACITY, "more than 2^25 points in one keyframe window");
    if (write_cloud) CU(ctx->s_cloud.ensure(n_ub * sizeof(PointXYZI)));
    DeskewArgs da{};
    {
        int r = size_chunks(ctx, da, n_ub, 1);
        if (r) return r;
    }
    if (!uploaded) {
        char* pin = reinterpret_cast<char*>(ctx->pinned);
        memcpy(pin + ctx->pin_ins(), ins, n_ins * sizeof(InsSample));
        CU(cudaMemcpyAsync(sp.ins, pin + ctx->pin_ins(), n_ins * sizeof(InsSample), cudaMemcpyHostToDevice, ctx->stream));
        KfState kf{};
        kf.ins_count = n_ins;
        kf.q_count = n_queries; // levins_keyframe: the features follow in the same call
        kf.stamp = stamp;
        memcpy(ctx->pinned, &kf, sizeof(kf));
        CU(cudaMemcpyAsync(sp.kf, ctx->pinned, sizeof(kf), cudaMemcpyHostToDevice, ctx->stream));
    }
    PrepareArgs pa{};
    pa.kfs = sp.kf;
    pa.pts = ctx->ring.as<PointXYZIT>();
    pa.pts_cap = ctx->cfg.max_points;
    pa.ins = sp.ins;
    pa.table = ctx->s_table.as<Interval>();
    pa.chunks = da.chunks;
    pa.chunks_per_kf = da.chunks_per_kf;
    pa.ring = ctx->ring_state.as<RingState>();
    pa.ext = ctx->ext;
    pa.c = ctx->consts;
    k_prepare<<<prepare_grid(ctx, 1, da.chunks_per_kf), 256, 0, ctx->stream>>>(pa);
    ctx->launches++;
    CU(cudaGetLastError());

    {
        int r = next_gen(ctx, ctx->s_zbuf.as<unsigned long long>(), CELLS, &ctx->s_gen);
        if (r) return r;
    }
    da.kfs = sp.kf;
    da.pts = ctx->ring.as<PointXYZIT>();
    da.pts_cap = ctx->cfg.max_points;
    da.table = ctx->s_table.as<Interval>();
    da.zbuf = ctx->s_zbuf.as<unsigned long long>();
    da.cloud_out = write_cloud ? ctx->s_cloud.as<PointXYZI>() : nullptr;
    da.gen = ctx->s_gen;
    da.ext = ctx->ext;
    da.c = ctx->consts;
    return launch_deskew(ctx, da, write_cloud);
}

int run_finalize_stream(levins_ctx* ctx, bool from_cloud, const PointXYZI* cloud, bool want_winner) {
    FinalizeArgs fa{};
    fa.kfs = ctx->s_kf.as<KfState>();
    fa.pts = ctx->ring.as<PointXYZIT>();
    fa.pts_cap = ctx->cfg.max_points;
    fa.table = ctx->s_table.as<Interval>();
    fa.cloud = cloud;
    fa.zbuf = ctx->s_zbuf.as<unsigned long long>();
    fa.cells = ctx->s_cells.as<float4>();
    fa.winner = want_winner ? ctx->s_winner.as<int32_t>() : nullptr;
    fa.gen = ctx->s_gen;
    fa.ext = ctx->ext;
    fa.c = ctx->consts;
    dim3 grid((CELLS + FINALIZE_THREADS - 1) / FINALIZE_THREADS, 1);
    if (from_cloud)
        k_finalize<true><<<grid, FINALIZE_THREADS, 0, ctx->stream>>>(fa);
    else
        k_finalize<false><<<grid, FINALIZE_THREADS, 0, ctx->stream>>>(fa);
    ctx->launches++;
    CU(cudaGetLastError());
    return LEVINS_OK;
}

// feature side on the device: tracked pixels in, accept gate + world points out (null camera: plain unit-ray queries)
struct PixelOpts {
    const levins_camera* cam = nullptr;
    int undistort = 0;
    const levins_pose* pose_ref = nullptr;
};

int run_associate_stream(levins_ctx* ctx, const StreamPtrs& sp, size_t nf, bool with_idx, bool fused, const PixelOpts& px = PixelOpts()) {
    AssocArgs aa{};
    if (px.cam) {
        aa.pixels = sp.pixels;
        aa.queries_out = sp.queries;
        aa.undist_out = sp.undist;
        aa.undistort = px.undistort;
        static_assert(sizeof(levins_camera) == sizeof(CameraModel), "camera layout");
        memcpy(&aa.cam, px.cam, sizeof(CameraModel));
    }
    aa.kfs = sp.kf;
    aa.queries = sp.queries;
    aa.cells = ctx->s_cells.as<float4>();
    aa.c2c = with_idx ? ctx->s_c2c.as<int32_t>() : nullptr;
    aa.zbuf = ctx->s_zbuf.as<unsigned long long>();
    aa.pts = ctx->ring.as<PointXYZIT>();
    aa.pts_cap = ctx->cfg.max_points;
    aa.table = ctx->s_table.as<Interval>();
    aa.gen = ctx->s_gen;
    aa.count_cells = 1; // one keyframe: counting M inside this launch is free
    aa.ext = ctx->ext;
    aa.nn = ctx->s_nn.as<float4>();
    aa.nn_meta = ctx->s_nn_meta.as<uint2>();
    aa.knn_idx = with_idx ? ctx->s_knn_idx.as<int32_t>() : nullptr;
    aa.knn_d2 = with_idx ? ctx->s_knn_d2.as<float>() : nullptr;
    aa.c = ctx->consts;
    dim3 grid((unsigned) ((nf + ASSOC_WARPS - 1) / ASSOC_WARPS), 1);
    if (fused)
        k_associate<true><<<grid, ASSOC_WARPS * 32, 0, ctx->stream>>>(aa);
    else
        k_associate<false><<<grid, ASSOC_WARPS * 32, 0, ctx->stream>>>(aa);
    ctx->launches++;
    CU(cudaGetLastError());
    PlaneArgs pa{};
    pa.queries = sp.queries;
    pa.nn = ctx->s_nn.as<float4>();
    pa.nn_meta = ctx->s_nn_meta.as<uint2>();
    pa.q_begin = 0;
    pa.q_count = nf;
    pa.type = sp.type;
    pa.xyz_depth = sp.xyzd;
    pa.flags = ctx->s_flags.as<int32_t>();
    pa.c = ctx->consts;
    if (sp.kf_out) {
        pa.kf_src = sp.kf;
        pa.kf_out = sp.kf_out;
        pa.ring_src = ctx->ring_state.as<RingState>();
        pa.ring_out = sp.ring_out;
    }
    if (px.cam) {
        pa.accepted = sp.accepted;
        if (px.pose_ref) {
            pa.pw = sp.pw;
            memcpy(pa.poseR, px.pose_ref->R, sizeof(pa.poseR));
            memcpy(pa.poset, px.pose_ref->t, sizeof(pa.poset));
        }
    }
    k_plane<<<(unsigned) ((nf + 127) / 128), 128, 0, ctx->stream>>>(pa);
    ctx->launches++;
    CU(cudaGetLastError());
    return LEVINS_OK;
}

} // namespace

// =================================================================================================
// lifecycle
// =================================================================================================
extern "C" int levins_abi_version(void) { return LEVINS_ABI_VERSION; }

extern "C" void levins_config_default(levins_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(levins_config);
    cfg->libm_mode = 0;
    cfg->stack_interval = 0.5;          // le_vins.h:130
    cfg->reserved_buffer_length = 10.0; // le_vins.h RESERVED_BUFFER_LENGTH
    cfg->nearest_depth = 1.0;           // visual_mappoint.h:50
    cfg->farthest_depth = 200.0;        // visual_mappoint.h:51
    cfg->plane_residual_threshold = 0.1;
    cfg->range_spread_threshold = 0.5;
    cfg->nearest_distance_counts = 3;
    cfg->max_points = 1ull << 24; // 16 Mi points = 512 MiB: 11 s of an OS1-128
    cfg->max_ins = 4096;
    cfg->max_features = 4096;
}

extern "C" const char* levins_status_string(int s) {
    switch (s) {
    case LEVINS_OK: return "ok";
    case LEVINS_EMPTY: return "empty";
    case LEVINS_ERR_ARG: return "invalid argument";
    case LEVINS_ERR_CUDA: return "CUDA error";
    case LEVINS_ERR_CAPACITY: return "capacity exceeded";
    case LEVINS_ERR_STATE: return "invalid state";
    default: return "unknown";
    }
}

extern "C" int levins_create(int device, const levins_config* cfg, levins_ctx** out) {
    if (!out) return LEVINS_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
        fprintf(stderr, "levins_create: no usable CUDA device %d (%s); this library has no CPU fallback\n", device,
                e == cudaSuccess ? "device count is 0 or ordinal out of range" : cudaGetErrorString(e));
        return LEVINS_ERR_CUDA;
    }
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        fprintf(stderr, "levins_create: device %d is sm_%d%d; the kernels are built for sm_100a only\n", device, prop.major,
                prop.minor);
        return LEVINS_ERR_CUDA;
    }
    levins_ctx* ctx = new levins_ctx();
    ctx->device = device;
    levins_config_default(&ctx->cfg);
    if (cfg) {
        size_t n = std::min<size_t>(cfg->struct_size ? cfg->struct_size : sizeof(levins_config), sizeof(levins_config));
        memcpy(&ctx->cfg, cfg, n);
        ctx->cfg.struct_size = sizeof(levins_config);
    }
    if (ctx->cfg.max_points == 0 || ctx->cfg.max_ins == 0 || ctx->cfg.stack_interval <= 0 || ctx->cfg.deskew_precision != 0) {
        delete ctx;
        return LEVINS_ERR_ARG;
    }
    ctx->consts = make_consts(ctx->cfg);
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost(&ctx->pinned, std::max<size_t>(ctx->pin_total(), 1 << 20)) != cudaSuccess) {
        delete ctx;
        return LEVINS_ERR_CUDA;
    }
    ctx->pinned_bytes = std::max<size_t>(ctx->pin_total(), 1 << 20);
    ctx->own_stream = true;
    *out = ctx;
    return LEVINS_OK;
}

extern "C" void levins_destroy(levins_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (DevBuf* b : {&ctx->ring, &ctx->ring_state, &ctx->chunks, &ctx->raw_stage, &ctx->raw_blocks, &ctx->raw_misc, &ctx->s_ins, &ctx->s_table, &ctx->s_kf, &ctx->s_queries, &ctx->s_type,
                      &ctx->s_xyzd, &ctx->s_flags, &ctx->s_knn_idx, &ctx->s_knn_d2, &ctx->s_nn, &ctx->s_nn_meta, &ctx->b_nn, &ctx->b_nn_meta, &ctx->s_cloud, &ctx->s_in_cloud,
                      &ctx->s_zbuf, &ctx->s_cells, &ctx->s_c2c, &ctx->s_unit, &ctx->s_winner, &ctx->b_pts, &ctx->b_ins,
                      &ctx->b_table, &ctx->b_queries, &ctx->b_kfs, &ctx->b_results, &ctx->b_flags, &ctx->b_zbuf,
                      &ctx->b_cells, &ctx->g_block, &ctx->b_packed, &ctx->b_scans, &ctx->s_pixels, &ctx->s_undist, &ctx->s_accepted, &ctx->s_pw, &ctx->scratch, &ctx->s_arena_in, &ctx->s_arena_out, &ctx->cull_stat})
        b->release();
    for (cudaEvent_t ev : ctx->ev_pool) cudaEventDestroy(ev);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->ev_binned) cudaEventDestroy(ctx->ev_binned);
    if (ctx->ev_counted) cudaEventDestroy(ctx->ev_counted);
    if (ctx->ev_append) cudaEventDestroy(ctx->ev_append);
    for (int r = 0; r < MAX_PEERS; r++)
        if (ctx->peer_mapped[r]) cudaIpcCloseMemHandle(ctx->peer_mapped[r]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->pin_io) cudaFreeHost(ctx->pin_io);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* levins_last_error(const levins_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int levins_set_stream(levins_ctx* ctx, void* cuda_stream) {
    if (!ctx) return LEVINS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
    return LEVINS_OK;
}

extern "C" int levins_synchronize(levins_ctx* ctx) {
    if (!ctx) return LEVINS_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return LEVINS_OK;
}

extern "C" int levins_set_extrinsics(levins_ctx* ctx, const levins_pose* pose_b_c, const levins_pose* pose_c_l, double td_b_l) {
    if (!ctx || !pose_b_c || !pose_c_l) return fail(ctx, LEVINS_ERR_ARG, "null pose");
    std::lock_guard<std::mutex> lk(ctx->mu);
    memcpy(ctx->ext.Rbc, pose_b_c->R, sizeof(double) * 9);
    memcpy(ctx->ext.tbc, pose_b_c->t, sizeof(double) * 3);
    memcpy(ctx->ext.Rcl, pose_c_l->R, sizeof(double) * 9);
    memcpy(ctx->ext.tcl, pose_c_l->t, sizeof(double) * 3);
    ctx->ext.td_b_l = td_b_l;
    ctx->have_ext = true;
    return LEVINS_OK;
}

extern "C" int levins_set_profiling(levins_ctx* ctx, int enabled) {
    if (!ctx) return LEVINS_ERR_ARG;
    ctx->profiling = enabled != 0;
    return LEVINS_OK;
}

extern "C" uint64_t levins_launch_count(const levins_ctx* ctx) { return ctx ? ctx->launches : 0; }

// =================================================================================================
// single-shot pose lookups (host): the same functions k_prepare runs on the device (device_math.cuh)
// =================================================================================================
extern "C" size_t levins_ins_window_index(const levins_ins_sample* w, size_t n, double time) {
    if (!w || n < 2 || w[0].time > time || w[n - 1].time <= time) return 0; // misc.cc:28-30
    // near-uniform sampling: guess from the mean spacing, fix up by bisection over what is left
    size_t lo = 1, hi = n - 1;
    size_t g = 1 + (size_t) ((time - w[0].time) * ((double) (n - 1) / (w[n - 1].time - w[0].time)));
    g = std::min(std::max(g, lo), hi);
    for (int it = 0; it < 80 && lo <= hi; it++) {
        if (time < w[g - 1].time) {
            hi = g - 1;
            g = it < 2 ? g - 1 : (lo + hi) / 2;
        } else if (time >= w[g].time) {
            lo = g + 1;
            g = it < 2 ? g + 1 : (lo + hi) / 2;
        } else {
            return g;
        }
        if (lo > hi) break;
        g = std::min(std::max(g, lo), hi);
    }
    return 0;
}

extern "C" void levins_state_interpolate(const levins_ins_sample* s0, const levins_ins_sample* s1, double midtime, levins_ins_sample* out) {
    if (!s0 || !s1 || !out) return;
    static_assert(sizeof(levins_ins_sample) == sizeof(InsSample), "INS sample layout");
    interpolate_state(*reinterpret_cast<const InsSample*>(s0), *reinterpret_cast<const InsSample*>(s1), midtime, out->p, out->q);
    out->time = midtime;
    out->state_time = midtime;
}

extern "C" void levins_state_to_pose(const levins_ins_sample* state, const levins_pose* ext, levins_pose* out) {
    if (!state || !ext || !out) return;
    pose_with_extrinsic(state->p, state->q, ext->R, ext->t, out->R, out->t);
}

extern "C" int levins_pose_at(const levins_ins_sample* w, size_t n, const levins_pose* ext, double time, levins_pose* out) {
    if (!w || !n || !ext || !out) return 0;
    const size_t i = levins_ins_window_index(w, n, time);
    if (i > 0) {
        interpolate_pose(*reinterpret_cast<const InsSample*>(w + i - 1), *reinterpret_cast<const InsSample*>(w + i), time, ext->R, ext->t,
                         out->R, out->t);
        return 1;
    }
    state_pose(*reinterpret_cast<const InsSample*>(w + n - 1), ext->R, ext->t, out->R, out->t); // misc.cc:74-77
    return 0;
}

// =================================================================================================
// streaming API
// =================================================================================================
extern "C" int levins_append_scan(levins_ctx* ctx, const levins_point_xyzit* points, size_t n) {
    if (!ctx || (!points && n)) return fail(ctx, LEVINS_ERR_ARG, "null points");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return LEVINS_OK;
    int r = ensure_ring(ctx);
    if (r) return r;
    const uint64_t cap = ctx->cfg.max_points;
    if (ctx->ring_count_ub + n > cap) { // learn the exact live count before refusing
        RingState st;
        r = read_ring_state(ctx, &st);
        if (r) return r;
        if (st.count + n > cap) return fail(ctx, LEVINS_ERR_CAPACITY, "ring buffer full: raise config.max_points");
    }
    if (!ctx->ring_tail_known) { // levins_append_raw moved the tail by a count only the device knows
        RingState st;
        r = read_ring_state(ctx, &st);
        if (r) return r;
        ctx->ring_tail = (st.head + st.count) % cap;
        ctx->ring_tail_known = true;
    }
    const uint64_t tail = ctx->ring_tail;
    const uint64_t first = std::min<uint64_t>(n, cap - tail);
    PointXYZIT* base = ctx->ring.as<PointXYZIT>();
    // (a scan in pageable memory is staged by the driver at ~10 GB/s and blocks the caller meanwhile; staging it here
    // into page-locked slots was measured and is no faster -- the host copy is the cost.  Scans written directly into
    // levins_alloc_host() buffers avoid it: ~10 us instead of ~80 us per 24 000-point scan.)
    CU(cudaMemcpyAsync(base + tail, points, first * sizeof(PointXYZIT), cudaMemcpyHostToDevice, ctx->stream));
    if (first < n)
        CU(cudaMemcpyAsync(base, points + first, (n - first) * sizeof(PointXYZIT), cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->ev_append) CU(cudaEventCreateWithFlags(&ctx->ev_append, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->ev_append, ctx->stream)); // levins_append_wait(): the caller's buffer has been read
    ctx->ring_tail = (tail + n) % cap;
    ctx->ring_count_ub += n;
    const double reserved = ctx->cfg.reserved_buffer_length + ctx->cfg.stack_interval * 2; // le_vins.cc:156
    k_ring_append<<<1, 32, 0, ctx->stream>>>(ctx->ring_state.as<RingState>(), base, cap, n, reserved);
    ctx->launches++;
    CU(cudaGetLastError());
    return LEVINS_OK;
}

extern "C" int levins_append_raw(levins_ctx* ctx, const levins_raw_format* fmt, const void* records, size_t n, uint64_t timebase_ns,
                                 double stamp, int to_gps_time, int scan_line, double nearest, double farthest, size_t* n_appended,
                                 double* t_start, double* t_end) {
    if (!ctx || !fmt || (!records && n)) return fail(ctx, LEVINS_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (n_appended) *n_appended = 0;
    if (t_start) *t_start = DBL_MAX; // lidar_converter.cc:50-51
    if (t_end) *t_end = 0;
    if ((fmt->kind != 1 && fmt->kind != 2) || fmt->point_step == 0 || (fmt->kind == 2 && (fmt->time_kind < 1 || fmt->time_kind > 3)))
        return fail(ctx, LEVINS_ERR_ARG, "unknown record format");
    const uint64_t need = fmt->kind == 1 ? 4u : (fmt->time_kind == 3 ? 8u : 4u);
    const uint64_t ends[] = {(uint64_t) fmt->off_x + 4, (uint64_t) fmt->off_y + 4, (uint64_t) fmt->off_z + 4,
                             (uint64_t) fmt->off_intensity + (fmt->kind == 1 ? 1u : 4u), (uint64_t) fmt->off_time + need,
                             fmt->kind == 1 ? (uint64_t) fmt->off_ring + 1 : 0u, fmt->kind == 1 ? (uint64_t) fmt->off_tag + 1 : 0u};
    for (uint64_t e : ends) // 64-bit sums: an offset near UINT32_MAX must not wrap past the check
        if (e > fmt->point_step) return fail(ctx, LEVINS_ERR_ARG, "field beyond point_step");
    if (n == 0) return LEVINS_OK;
    if (n > 0x7fffffffull) return fail(ctx, LEVINS_ERR_ARG, "scan too large");
    int r = ensure_ring(ctx);
    if (r) return r;
    const uint64_t cap = ctx->cfg.max_points;
    if (ctx->ring_count_ub + n > cap) { // learn the exact live count before refusing
        RingState st;
        r = read_ring_state(ctx, &st);
        if (r) return r;
        if (st.count + n > cap) return fail(ctx, LEVINS_ERR_CAPACITY, "ring buffer full: raise config.max_points");
    }
    const size_t bytes = n * (size_t) fmt->point_step;
    const unsigned nblocks = (unsigned) ((n + RAW_BLOCK - 1) / RAW_BLOCK);
    CU(ctx->raw_stage.ensure(bytes));
    CU(ctx->raw_blocks.ensure((size_t) nblocks * 4));
    char* pin = reinterpret_cast<char*>(ctx->pinned);
    if (!ctx->raw_misc.p) { // {min, max} keys | result | ticket; the kernels re-arm the keys after every scan
        CU(ctx->raw_misc.ensure(64));
        const double init[2] = {DBL_MAX, 0.0}; // start = DBL_MAX, end = 0 as ordered keys (positive doubles: sign bit set)
        unsigned long long keys[8] = {};
        for (int i = 0; i < 2; i++) {
            unsigned long long b;
            memcpy(&b, &init[i], 8);
            keys[i] = b | 0x8000000000000000ull;
        }
        memcpy(pin + 512, keys, 64);
        CU(cudaMemcpyAsync(ctx->raw_misc.p, pin + 512, 64, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaMemcpyAsync(ctx->raw_stage.p, records, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->ev_append) CU(cudaEventCreateWithFlags(&ctx->ev_append, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->ev_append, ctx->stream)); // levins_append_wait(): the caller's buffer has been read
    RawArgs a{};
    a.rec = ctx->raw_stage.as<unsigned char>();
    a.n = (unsigned) n;
    memcpy(&a.f, fmt, sizeof(RawFormat));
    a.timebase_ns = timebase_ns;
    a.stamp = stamp;
    if (fmt->kind == 2 && fmt->time_kind != 3 && to_gps_time) { // lidar_converter.cc:92-98, 141-147 (host libm like the reference)
        const double seconds = stamp + 18 - 315964800;
        const int week = (int) floor(seconds / 604800);
        a.stamp = seconds - week * 604800;
    }
    a.to_gps = to_gps_time ? 1 : 0;
    a.scan_line = scan_line;
    a.near2 = nearest * nearest; // :31-32
    a.far2 = farthest * farthest;
    a.block_kept = ctx->raw_blocks.as<unsigned>();
    a.minmax = ctx->raw_misc.as<unsigned long long>();
    a.result = reinterpret_cast<RawResult*>(ctx->raw_misc.as<unsigned char>() + 16);
    a.ticket = reinterpret_cast<unsigned*>(ctx->raw_misc.as<unsigned char>() + 48);
    a.ring = ctx->ring.as<PointXYZIT>();
    a.cap = cap;
    a.ring_state = ctx->ring_state.as<RingState>();
    a.reserved_time = ctx->cfg.reserved_buffer_length + ctx->cfg.stack_interval * 2; // le_vins.cc:156
    k_raw_count<<<nblocks, RAW_BLOCK, 0, ctx->stream>>>(a);
    k_raw_scatter<<<nblocks, RAW_BLOCK, 0, ctx->stream>>>(a); // + block offsets, result record, ring bookkeeping (last block)
    ctx->launches += 2;
    CU(cudaGetLastError());
    ctx->ring_count_ub += n;       // an upper bound until the next read-back
    ctx->ring_tail_known = false;  // the tail moved by a count only the device knows
    if (!n_appended && !t_start && !t_end) return LEVINS_OK; // nothing to report: fully asynchronous
    CU(cudaMemcpyAsync(pin + 640, a.result, sizeof(RawResult), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    RawResult res;
    memcpy(&res, pin + 640, sizeof(res));
    ctx->ring_count_ub -= n - res.count;
    if (n_appended) *n_appended = (size_t) res.count;
    if (t_start) *t_start = res.start;
    if (t_end) *t_end = res.end;
    return LEVINS_OK;
}

extern "C" int levins_append_wait(levins_ctx* ctx) {
    if (!ctx) return LEVINS_ERR_ARG;
    cudaEvent_t ev;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ev = ctx->ev_append;
    }
    if (!ev) return LEVINS_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ev));
    return LEVINS_OK;
}

extern "C" int levins_alloc_host(size_t bytes, void** out) {
    if (!out) return LEVINS_ERR_ARG;
    *out = nullptr;
    return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? LEVINS_OK : LEVINS_ERR_CUDA;
}
extern "C" int levins_free_host(void* p) { return (!p || cudaFreeHost(p) == cudaSuccess) ? LEVINS_OK : LEVINS_ERR_CUDA; }

extern "C" int levins_buffer_info(levins_ctx* ctx, size_t* n_points, double* t_front, double* t_back) {
    if (!ctx) return LEVINS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (n_points) *n_points = 0;
    if (t_front) *t_front = 0;
    if (t_back) *t_back = 0;
    if (!ctx->ring.p) return LEVINS_OK;
    RingState st;
    int r = read_ring_state(ctx, &st);
    if (r) return r;
    if (n_points) *n_points = st.count;
    if (st.count) {
        const uint64_t cap = ctx->cfg.max_points;
        PointXYZIT a, b;
        CU(cudaMemcpy(&a, ctx->ring.as<PointXYZIT>() + st.head, sizeof(a), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&b, ctx->ring.as<PointXYZIT>() + (st.head + st.count - 1) % cap, sizeof(b), cudaMemcpyDeviceToHost));
        if (t_front) *t_front = a.time;
        if (t_back) *t_back = b.time;
    }
    return LEVINS_OK;
}

extern "C" int levins_buffer_clear(levins_ctx* ctx) {
    if (!ctx) return LEVINS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (ctx->ring.p) CU(cudaMemsetAsync(ctx->ring_state.p, 0, sizeof(RingState), ctx->stream));
    ctx->ring_tail = 0;
    ctx->ring_tail_known = true;
    ctx->ring_count_ub = 0;
    ctx->s_cloud_valid = false;
    ctx->s_image_valid = false;
    return LEVINS_OK;
}

// read KfState + ring state back (one small copy, synchronises)
static int fetch_kf(levins_ctx* ctx, KfState* kf) {
    char* pin = reinterpret_cast<char*>(ctx->pinned);
    CU(cudaMemcpyAsync(pin, ctx->s_kf.p, sizeof(KfState), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(pin + 256, ctx->ring_state.p, sizeof(RingState), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(kf, pin, sizeof(KfState));
    RingState st;
    memcpy(&st, pin + 256, sizeof(st));
    ctx->ring_count_ub = st.count;
    return LEVINS_OK;
}

extern "C" int levins_stack_deskew(levins_ctx* ctx, double stamp, const levins_ins_sample* ins, size_t n_ins,
                                   levins_point_xyzi* out_cloud, size_t out_capacity, size_t* n_out) {
    if (!ctx) return LEVINS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (n_out) *n_out = 0;
    ctx->s_cloud_valid = false;
    ctx->s_image_valid = false;
    int r = ensure_stream_buffers(ctx, n_ins, 1);
    if (r) return r;
    r = run_stage1(ctx, default_ptrs(ctx), stamp, ins, n_ins, /*write_cloud=*/true);
    if (r) return r;
    KfState kf;
    r = fetch_kf(ctx, &kf);
    if (r) return r;
    ctx->s_cloud_n = kf.n_proj;
    ctx->s_cloud_valid = true;
    ctx->s_image_from_raw = true; // the z-buffer now holds this keyframe's winners (finalised lazily)
    if (n_out) *n_out = kf.n_proj;
    if (out_cloud && kf.n_proj) {
        if (out_capacity < kf.n_proj) return fail(ctx, LEVINS_ERR_CAPACITY, "out_cloud too small");
        CU(cudaMemcpyAsync(out_cloud, ctx->s_cloud.p, (size_t) kf.n_proj * sizeof(PointXYZI), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return kf.status ? LEVINS_EMPTY : LEVINS_OK;
}

extern "C" int levins_build_range_image(levins_ctx* ctx, const levins_point_xyzi* cloud, size_t n,
                                        levins_point_xyzi* unit_sphere_out, size_t unit_capacity, int32_t* cell_winner_out,
                                        size_t* m_out) {
    if (!ctx) return LEVINS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (m_out) *m_out = 0;
    int r = ensure_stream_buffers(ctx, 1, 1);
    if (r) return r;
    ctx->s_image_valid = false;
    const PointXYZI* dcloud = nullptr;
    if (cloud) { // explicit cloud: re-bin from scratch
        if (n > 0xffffffffull) return fail(ctx, LEVINS_ERR_ARG, "cloud too large");
        if (n >= MAX_POINTS_PER_KEYFRAME) return fail(ctx, LEVINS_ERR_CAPACITY, "more than 2^25 points in one cloud");
        ctx->s_image_from_raw = false;
        r = next_gen(ctx, ctx->s_zbuf.as<unsigned long long>(), CELLS, &ctx->s_gen); // a new, empty image
        if (r) return r;
        CU(ctx->s_in_cloud.ensure(std::max<size_t>(n, 1) * sizeof(PointXYZI)));
        if (n) CU(cudaMemcpyAsync(ctx->s_in_cloud.p, cloud, n * sizeof(PointXYZI), cudaMemcpyHostToDevice, ctx->stream));
        dcloud = ctx->s_in_cloud.as<PointXYZI>();
        KfState kf{};
        memcpy(ctx->pinned, &kf, sizeof(kf));
        CU(cudaMemcpyAsync(ctx->s_kf.p, ctx->pinned, sizeof(kf), cudaMemcpyHostToDevice, ctx->stream));
        if (n) {
            const unsigned grid = (unsigned) ((n + 255) / 256);
            if (ctx->consts.libm_mode == 0)
                k_bin_cloud<0><<<grid, 256, 0, ctx->stream>>>(dcloud, (unsigned) n, ctx->s_zbuf.as<unsigned long long>(), ctx->s_gen,
                                                                 ctx->consts);
            else
                k_bin_cloud<1><<<grid, 256, 0, ctx->stream>>>(dcloud, (unsigned) n, ctx->s_zbuf.as<unsigned long long>(), ctx->s_gen,
                                                                 ctx->consts);
            ctx->launches++;
            CU(cudaGetLastError());
        }
        r = run_finalize_stream(ctx, true, dcloud, true);
        if (r) return r;
    } else {
        if (!ctx->s_cloud_valid || !ctx->s_image_from_raw)
            return fail(ctx, LEVINS_ERR_STATE, "no resident cloud: call levins_stack_deskew first or pass a cloud");
        // the z-buffer was filled by the fused deskew+bin kernel; winners index the resident cloud
        r = run_finalize_stream(ctx, true, ctx->s_cloud.as<PointXYZI>(), true);
        if (r) return r;
        ctx->s_image_from_raw = false;
    }
    k_compact<<<1, 1024, 0, ctx->stream>>>(ctx->s_cells.as<float4>(), ctx->s_c2c.as<int32_t>(), ctx->s_unit.as<PointXYZI>());
    ctx->launches++;
    CU(cudaGetLastError());
    KfState kf;
    CU(cudaMemcpyAsync(ctx->pinned, ctx->s_kf.p, sizeof(KfState), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(&kf, ctx->pinned, sizeof(kf));
    const size_t m = kf.m;
    if (m_out) *m_out = m;
    if (unit_sphere_out && m) {
        if (unit_capacity < m) return fail(ctx, LEVINS_ERR_CAPACITY, "unit_sphere_out too small");
        CU(cudaMemcpyAsync(unit_sphere_out, ctx->s_unit.p, m * sizeof(PointXYZI), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (cell_winner_out)
        CU(cudaMemcpyAsync(cell_winner_out, ctx->s_winner.p, (size_t) CELLS * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->s_image_valid = m > 0;
    return m ? LEVINS_OK : LEVINS_EMPTY;
}

static int associate_body(levins_ctx* ctx, const float* query_xyz, const float* pixels, size_t nf, const PixelOpts& px, int8_t* type,
                          float* xyz_depth, int32_t* knn_idx, float* knn_d2, int32_t* flags, uint8_t* accepted, double* pw, float* undistorted) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (nf > ctx->cfg.max_features) return fail(ctx, LEVINS_ERR_CAPACITY, "more features than config.max_features");
    if (!ctx->s_image_valid) return fail(ctx, LEVINS_ERR_STATE, "no range image: call levins_build_range_image first");
    if (nf == 0) return LEVINS_OK;
    int r = ensure_stream_buffers(ctx, 1, nf);
    if (r) return r;
    if (pixels)
        CU(cudaMemcpyAsync(ctx->s_pixels.p, pixels, nf * 8, cudaMemcpyHostToDevice, ctx->stream));
    else
        CU(cudaMemcpyAsync(ctx->s_queries.p, query_xyz, nf * 12, cudaMemcpyHostToDevice, ctx->stream));
    // q_count lives in the device KfState
    unsigned long long qc[2] = {0, nf};
    memcpy(ctx->pinned, qc, sizeof(qc));
    CU(cudaMemcpyAsync(reinterpret_cast<char*>(ctx->s_kf.p) + offsetof(KfState, q_begin), ctx->pinned, sizeof(qc),
                       cudaMemcpyHostToDevice, ctx->stream));
    r = run_associate_stream(ctx, default_ptrs(ctx), nf, true, false, px);
    if (r) return r;
    if (type) CU(cudaMemcpyAsync(type, ctx->s_type.p, nf, cudaMemcpyDeviceToHost, ctx->stream));
    if (xyz_depth) CU(cudaMemcpyAsync(xyz_depth, ctx->s_xyzd.p, nf * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (flags) CU(cudaMemcpyAsync(flags, ctx->s_flags.p, nf * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (knn_idx) CU(cudaMemcpyAsync(knn_idx, ctx->s_knn_idx.p, nf * 20, cudaMemcpyDeviceToHost, ctx->stream));
    if (knn_d2) CU(cudaMemcpyAsync(knn_d2, ctx->s_knn_d2.p, nf * 20, cudaMemcpyDeviceToHost, ctx->stream));
    if (accepted) CU(cudaMemcpyAsync(accepted, ctx->s_accepted.p, nf, cudaMemcpyDeviceToHost, ctx->stream));
    if (pw && px.pose_ref) CU(cudaMemcpyAsync(pw, ctx->s_pw.p, nf * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (undistorted) CU(cudaMemcpyAsync(undistorted, ctx->s_undist.p, nf * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LEVINS_OK;
}

extern "C" int levins_associate(levins_ctx* ctx, const float* query_xyz, size_t nf, int8_t* type, float* xyz_depth,
                                int32_t* knn_idx, float* knn_d2, int32_t* flags) {
    if (!ctx || (!query_xyz && nf)) return fail(ctx, LEVINS_ERR_ARG, "null queries");
    return associate_body(ctx, query_xyz, nullptr, nf, PixelOpts(), type, xyz_depth, knn_idx, knn_d2, flags, nullptr, nullptr, nullptr);
}

extern "C" int levins_associate_pixels(levins_ctx* ctx, const levins_camera* camera, const float* pixels, size_t nf, int undistort,
                                       const levins_pose* pose_ref, int8_t* type, float* xyz_depth, uint8_t* accepted, double* pw,
                                       float* undistorted) {
    if (!ctx || !camera || (!pixels && nf)) return fail(ctx, LEVINS_ERR_ARG, "null camera or pixels");
    if (pw && !pose_ref) return fail(ctx, LEVINS_ERR_ARG, "world points need pose_ref");
    PixelOpts px;
    px.cam = camera;
    px.undistort = undistort;
    px.pose_ref = pose_ref;
    return associate_body(ctx, nullptr, pixels, nf, px, type, xyz_depth, nullptr, nullptr, nullptr, accepted, pw, undistorted);
}

// The fused keyframe.  Everything the call sends travels in ONE host->device copy (KfState | INS window | rays or pixels, packed
// in page-locked memory) and everything it returns in ONE device->host copy (final KfState | RingState | type | xyz_depth |
// accepted | undistorted | world points: k_plane leaves the states next to the results).
static int keyframe_body(levins_ctx* ctx, double stamp, const levins_ins_sample* ins, size_t n_ins, const float* query_xyz,
                         const float* pixels, size_t nf, const PixelOpts& px, int8_t* type, float* xyz_depth, uint8_t* accepted, double* pw,
                         float* undistorted, size_t* n_projected_out, size_t* m_out) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (nf > ctx->cfg.max_features) return fail(ctx, LEVINS_ERR_CAPACITY, "more features than config.max_features");
    if (n_ins > ctx->cfg.max_ins) return fail(ctx, LEVINS_ERR_CAPACITY, "INS window larger than config.max_ins");
    if (n_projected_out) *n_projected_out = 0;
    if (m_out) *m_out = 0;
    ctx->s_cloud_valid = false;
    ctx->s_image_valid = false;
    int r = ensure_stream_buffers(ctx, n_ins, nf);
    if (r) return r;
    ctx->s_image_from_raw = false;
    auto a16 = [](size_t v) { return (v + 15) & ~size_t(15); };
    // input arena
    const size_t in_ins = 256, in_q = in_ins + a16(n_ins * sizeof(InsSample)), in_px = in_q + a16(nf * 12), in_end = in_px + a16(nf * 8);
    // output arena
    const size_t o_ring = 128, o_type = 256, o_xyzd = o_type + a16(nf), o_acc = o_xyzd + nf * 16, o_und = o_acc + a16(nf),
                 o_pw = o_und + a16(nf * 8), o_end = o_pw + nf * 24;
    CU(ctx->s_arena_in.ensure(in_end));
    CU(ctx->s_arena_out.ensure(o_end));
    if (ctx->pin_io_bytes < in_end + o_end) {
        if (ctx->pin_io) cudaFreeHost(ctx->pin_io);
        ctx->pin_io = nullptr;
        ctx->pin_io_bytes = 0;
        const size_t want = (in_end + o_end) * 2 + 4096;
        CU(cudaMallocHost(&ctx->pin_io, want));
        ctx->pin_io_bytes = want;
    }
    char* hin = static_cast<char*>(ctx->pin_io);
    char* hout = hin + ctx->pin_io_bytes / 2;
    char* din = ctx->s_arena_in.as<char>();
    char* dout = ctx->s_arena_out.as<char>();
    StreamPtrs sp{};
    sp.kf = reinterpret_cast<KfState*>(din);
    sp.ins = reinterpret_cast<InsSample*>(din + in_ins);
    sp.queries = reinterpret_cast<float*>(din + in_q);
    sp.pixels = reinterpret_cast<float*>(din + in_px);
    sp.type = reinterpret_cast<int8_t*>(dout + o_type);
    sp.xyzd = reinterpret_cast<float4*>(dout + o_xyzd);
    sp.accepted = reinterpret_cast<unsigned char*>(dout + o_acc);
    sp.undist = reinterpret_cast<float*>(dout + o_und);
    sp.pw = reinterpret_cast<double*>(dout + o_pw);
    sp.kf_out = reinterpret_cast<KfState*>(dout);
    sp.ring_out = reinterpret_cast<RingState*>(dout + o_ring);

    if (!ctx->have_ext) return fail(ctx, LEVINS_ERR_STATE, "levins_set_extrinsics has not been called");
    if (!ins || n_ins == 0) return fail(ctx, LEVINS_ERR_ARG, "empty INS window");
    const bool empty = !ctx->ring.p || ctx->ring_count_ub == 0;
    if (!empty) {
        KfState kf{};
        kf.ins_count = n_ins;
        kf.q_count = nf;
        kf.stamp = stamp;
        static_assert(sizeof(KfState) <= 256, "input arena header");
        memcpy(hin, &kf, sizeof(kf));
        memcpy(hin + in_ins, ins, n_ins * sizeof(InsSample));
        size_t used = in_q;
        if (nf && pixels) {
            memcpy(hin + in_px, pixels, nf * 8);
            used = in_px + nf * 8;
        } else if (nf) {
            memcpy(hin + in_q, query_xyz, nf * 12);
            used = in_q + nf * 12;
        }
        CU(cudaMemcpyAsync(din, hin, used, cudaMemcpyHostToDevice, ctx->stream));
    }
    r = empty ? LEVINS_EMPTY : run_stage1(ctx, sp, stamp, ins, n_ins, /*write_cloud=*/false, nf, /*uploaded=*/true);
    if (r == LEVINS_EMPTY) {
        for (size_t k = 0; k < nf; k++) {
            if (type) type[k] = LEVINS_MAPPOINT_NONE;
            if (accepted) accepted[k] = 0;
            if (xyz_depth && query_xyz) {
                memcpy(xyz_depth + 4 * k, query_xyz + 3 * k, 12);
                xyz_depth[4 * k + 3] = 0;
            }
        }
        return LEVINS_EMPTY;
    }
    if (r) return r;
    const bool want_pw = pw && px.pose_ref;
    KfState kf;
    if (nf) {
        r = run_associate_stream(ctx, sp, nf, false, true, px); // also counts M; k_plane leaves the states in the output arena
        if (r) return r;
        size_t span = o_xyzd + nf * 16; // KfState | RingState | type | xyz_depth are always fetched
        if (px.cam) span = want_pw ? o_end : (undistorted ? o_pw : o_und);
        CU(cudaMemcpyAsync(hout, dout, span, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream)); // the one synchronisation of the call
        memcpy(&kf, hout, sizeof(KfState));
        RingState st;
        memcpy(&st, hout + o_ring, sizeof(st));
        ctx->ring_count_ub = st.count;
        if (type) memcpy(type, hout + o_type, nf);
        if (xyz_depth) memcpy(xyz_depth, hout + o_xyzd, nf * 16);
        if (accepted && px.cam) memcpy(accepted, hout + o_acc, nf);
        if (undistorted && px.cam) memcpy(undistorted, hout + o_und, nf * 8);
        if (want_pw) memcpy(pw, hout + o_pw, nf * 24);
    } else { // no feature kernel to piggy-back the cell count and the state copy on
        k_count_cells<<<dim3(64, 1), 256, 0, ctx->stream>>>(sp.kf, ctx->s_zbuf.as<unsigned long long>(), ctx->s_gen);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hout, sp.kf, sizeof(KfState), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(hout + o_ring, ctx->ring_state.p, sizeof(RingState), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        memcpy(&kf, hout, sizeof(KfState));
        RingState st;
        memcpy(&st, hout + o_ring, sizeof(st));
        ctx->ring_count_ub = st.count;
    }
    if (n_projected_out) *n_projected_out = kf.n_proj;
    if (m_out) *m_out = kf.m;
    // the fused path keeps no cell array: a later levins_associate() needs levins_build_range_image() first
    ctx->s_image_valid = false;
    if ((++ctx->kf_calls & 31u) == 0) { // now and then: does the frustum cull ever fire for this sensor?
        r = update_cull_policy(ctx);
        if (r) return r;
    }
    return (kf.status || kf.m == 0) ? LEVINS_EMPTY : LEVINS_OK;
}

extern "C" int levins_keyframe(levins_ctx* ctx, double stamp, const levins_ins_sample* ins, size_t n_ins,
                               const float* query_xyz, size_t nf, int8_t* type, float* xyz_depth, size_t* n_projected_out,
                               size_t* m_out) {
    if (!ctx || (!query_xyz && nf)) return fail(ctx, LEVINS_ERR_ARG, "null queries");
    return keyframe_body(ctx, stamp, ins, n_ins, query_xyz, nullptr, nf, PixelOpts(), type, xyz_depth, nullptr, nullptr, nullptr,
                         n_projected_out, m_out);
}

extern "C" int levins_keyframe_pixels(levins_ctx* ctx, double stamp, const levins_ins_sample* ins, size_t n_ins,
                                      const levins_camera* camera, const float* pixels, size_t nf, int undistort,
                                      const levins_pose* pose_ref, int8_t* type, float* xyz_depth, uint8_t* accepted, double* pw,
                                      float* undistorted, size_t* n_projected_out, size_t* m_out) {
    if (!ctx || !camera || (!pixels && nf)) return fail(ctx, LEVINS_ERR_ARG, "null camera or pixels");
    if (pw && !pose_ref) return fail(ctx, LEVINS_ERR_ARG, "world points need pose_ref");
    PixelOpts px;
    px.cam = camera;
    px.undistort = undistort;
    px.pose_ref = pose_ref;
    return keyframe_body(ctx, stamp, ins, n_ins, nullptr, pixels, nf, px, type, xyz_depth, accepted, pw, undistorted, n_projected_out,
                         m_out);
}

extern "C" int levins_point_cloud_projection(levins_ctx* ctx, const levins_pose* pose, const levins_point_xyzi* src,
                                             levins_point_xyzi* dst, size_t n) {
    if (!ctx || !pose || (n && (!src || !dst))) return fail(ctx, LEVINS_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return LEVINS_OK;
    if (n > 0xffffffffull) return fail(ctx, LEVINS_ERR_ARG, "cloud too large");
    CU(ctx->s_in_cloud.ensure(2 * n * sizeof(PointXYZI)));
    PointXYZI* d_src = ctx->s_in_cloud.as<PointXYZI>();
    PointXYZI* d_dst = d_src + n;
    CU(cudaMemcpyAsync(d_src, src, n * sizeof(PointXYZI), cudaMemcpyHostToDevice, ctx->stream));
    Extrinsics e{};
    memcpy(e.Rcl, pose->R, sizeof(double) * 9);
    memcpy(e.tcl, pose->t, sizeof(double) * 3);
    k_project<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(d_src, d_dst, (unsigned) n, e);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dst, d_dst, n * sizeof(PointXYZI), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LEVINS_OK;
}

// =================================================================================================
// batched plane geometry and depth-factor / chi-square batches (rank 4)
// =================================================================================================
namespace {
// one scratch allocation carved into 256-byte aligned pieces
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T> T* take(size_t n) {
        T* r = reinterpret_cast<T*>(base + off);
        off += (n * sizeof(T) + 255) & ~size_t(255);
        return r;
    }
    static size_t need(std::initializer_list<size_t> bytes) {
        size_t t = 0;
        for (size_t b : bytes) t += (b + 255) & ~size_t(255);
        return t;
    }
};
} // namespace

extern "C" int levins_plane_estimation_batch(levins_ctx* ctx, const levins_point_xyzi* points, size_t n_groups, double threshold,
                                             uint8_t* valid, double* unit_normal, double* norm_inverse, double* distance,
                                             int32_t* n_distance) {
    if (!ctx || (!points && n_groups)) return fail(ctx, LEVINS_ERR_ARG, "null points");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (n_groups == 0) return LEVINS_OK;
    if (n_groups > 0x7fffffffull / 5) return fail(ctx, LEVINS_ERR_ARG, "too many groups");
    const size_t n = n_groups;
    CU(ctx->scratch.ensure(Carver::need({n * 5 * sizeof(PointXYZI), n, n * 24, n * 8, n * 40, n * 4})));
    Carver c(ctx->scratch.p);
    PointXYZI* d_pts = c.take<PointXYZI>(n * 5);
    unsigned char* d_valid = c.take<unsigned char>(n);
    double* d_un = c.take<double>(n * 3);
    double* d_ninv = c.take<double>(n);
    double* d_dist = c.take<double>(n * 5);
    int* d_nd = c.take<int>(n);
    CU(cudaMemcpyAsync(d_pts, points, n * 5 * sizeof(PointXYZI), cudaMemcpyHostToDevice, ctx->stream));
    k_plane_estimation<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(d_pts, (unsigned) n, threshold, d_valid, d_un, d_ninv, d_dist, d_nd);
    ctx->launches++;
    CU(cudaGetLastError());
    if (valid) CU(cudaMemcpyAsync(valid, d_valid, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (unit_normal) CU(cudaMemcpyAsync(unit_normal, d_un, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (norm_inverse) CU(cudaMemcpyAsync(norm_inverse, d_ninv, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (distance) CU(cudaMemcpyAsync(distance, d_dist, n * 40, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_distance) CU(cudaMemcpyAsync(n_distance, d_nd, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LEVINS_OK;
}

extern "C" int levins_point_to_plane_std_batch(levins_ctx* ctx, const levins_point_xyzi* nearest, size_t group_size,
                                               const levins_point_xyzi* targets, size_t n_groups, double point_std, double* std_out) {
    if (!ctx || !std_out || ((!nearest || !targets) && n_groups) || group_size == 0) return fail(ctx, LEVINS_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (n_groups == 0) return LEVINS_OK;
    if (n_groups > 0x7fffffffull || group_size > 0xffffu) return fail(ctx, LEVINS_ERR_ARG, "too many groups");
    const size_t n = n_groups;
    CU(ctx->scratch.ensure(Carver::need({n * group_size * sizeof(PointXYZI), n * sizeof(PointXYZI), n * 8})));
    Carver c(ctx->scratch.p);
    PointXYZI* d_near = c.take<PointXYZI>(n * group_size);
    PointXYZI* d_tgt = c.take<PointXYZI>(n);
    double* d_out = c.take<double>(n);
    CU(cudaMemcpyAsync(d_near, nearest, n * group_size * sizeof(PointXYZI), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_tgt, targets, n * sizeof(PointXYZI), cudaMemcpyHostToDevice, ctx->stream));
    k_point_to_plane_std<<<(unsigned) ((n + 127) / 128), 128, 0, ctx->stream>>>(d_near, (unsigned) group_size, d_tgt, (unsigned) n, point_std, d_out);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(std_out, d_out, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LEVINS_OK;
}

extern "C" int levins_depth_factor_batch(levins_ctx* ctx, const double* depth, const double* invdepth, size_t n, double sigma,
                                         double huber_a, double* residual, double* jacobian, double* cost) {
    if (!ctx || ((!depth || !invdepth) && n) || !(sigma > 0)) return fail(ctx, LEVINS_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return LEVINS_OK;
    if (n > 0x7fffffffull) return fail(ctx, LEVINS_ERR_ARG, "too many factors");
    CU(ctx->scratch.ensure(Carver::need({n * 8, n * 8, n * 8, n * 8, n * 8})));
    Carver c(ctx->scratch.p);
    double *d_depth = c.take<double>(n), *d_inv = c.take<double>(n), *d_r = c.take<double>(n), *d_j = c.take<double>(n), *d_c = c.take<double>(n);
    CU(cudaMemcpyAsync(d_depth, depth, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_inv, invdepth, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_depth_factor<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(d_depth, d_inv, (unsigned) n, sigma, huber_a, d_r, d_j, d_c);
    ctx->launches++;
    CU(cudaGetLastError());
    if (residual) CU(cudaMemcpyAsync(residual, d_r, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (jacobian) CU(cudaMemcpyAsync(jacobian, d_j, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (cost) CU(cudaMemcpyAsync(cost, d_c, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LEVINS_OK;
}

extern "C" int levins_chi2_culling_batch(levins_ctx* ctx, const double* residuals, size_t n, size_t dim, double threshold, double* cost,
                                         uint8_t* outlier) {
    if (!ctx || !outlier || (!residuals && n) || dim == 0) return fail(ctx, LEVINS_ERR_ARG, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return LEVINS_OK;
    if (n > 0x7fffffffull || dim > 64) return fail(ctx, LEVINS_ERR_ARG, "too many residual blocks or dimensions");
    CU(ctx->scratch.ensure(Carver::need({n * dim * 8, n * 8, n})));
    Carver c(ctx->scratch.p);
    double* d_r = c.take<double>(n * dim);
    double* d_c = c.take<double>(n);
    unsigned char* d_o = c.take<unsigned char>(n);
    CU(cudaMemcpyAsync(d_r, residuals, n * dim * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_chi2<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(d_r, (unsigned) n, (unsigned) dim, threshold, d_c, d_o);
    ctx->launches++;
    CU(cudaGetLastError());
    if (cost) CU(cudaMemcpyAsync(cost, d_c, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(outlier, d_o, n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LEVINS_OK;
}

// =================================================================================================
// batch API
// =================================================================================================
extern "C" int levins_batch_set_wave(levins_ctx* ctx, int w) {
    if (!ctx || w < 0 || w > 4096) return LEVINS_ERR_ARG;
    ctx->b_wave = w ? w : 256;
    return LEVINS_OK;
}

extern "C" size_t levins_result_slot_bytes(size_t n_queries) {
    const size_t nq16 = (std::max<size_t>(n_queries, 1) + 15) & ~size_t(15);
    return 17 * nq16;
}

// shared by the two upload entry points: validate the descriptors, size the pools, copy INS / queries / keyframe states.
// The point pool itself is filled by the caller (a plain copy, or a copy of packed records + k_unpack).
static int batch_prepare(levins_ctx* ctx, size_t n_points, const levins_ins_sample* ins, size_t n_ins, const float* query_xyz,
                         size_t n_queries, const levins_keyframe_desc* descs, size_t n_kf, bool borrowed = false) {
    ctx->b_valid = false;
    ctx->b_pts_view = nullptr;
    ctx->b_ins_view = nullptr;
    ctx->b_queries_view = nullptr;
    if (n_kf == 0) return fail(ctx, LEVINS_ERR_ARG, "no keyframes");
    std::vector<KfState>& kfs = ctx->b_kfs_host; // a member: the copy below stays asynchronous
    kfs.resize(n_kf);
    size_t table_entries = 0; // the table depends on the keyframe pose, so INS slices shared by keyframes get separate tables
    for (size_t k = 0; k < n_kf; k++) {
        const levins_keyframe_desc& d = descs[k];
        if (d.pt_begin + d.pt_count > n_points || d.ins_begin + d.ins_count > n_ins || d.q_begin + d.q_count > n_queries ||
            d.ins_count == 0 || d.pt_count >= MAX_POINTS_PER_KEYFRAME)
            return fail(ctx, LEVINS_ERR_ARG, "keyframe descriptor out of range");
        KfState s{};
        s.pt_begin = d.pt_begin;
        s.pt_count = d.pt_count;
        s.ins_begin = d.ins_begin;
        s.ins_count = d.ins_count;
        s.q_begin = d.q_begin;
        s.q_count = d.q_count;
        s.stamp = d.stamp;
        s.table_begin = table_entries;
        table_entries += d.ins_count + 1; // + the fallback copy (k_prepare)
        kfs[k] = s;
    }
    if (!borrowed) {
        CU(ctx->b_pts.ensure(std::max<size_t>(n_points, 1) * sizeof(PointXYZIT)));
        CU(ctx->b_ins.ensure(n_ins * sizeof(InsSample)));
        CU(ctx->b_queries.ensure(std::max<size_t>(n_queries, 1) * 12));
    }
    CU(ctx->b_table.ensure(table_entries * sizeof(Interval)));
    CU(ctx->b_kfs.ensure(n_kf * sizeof(KfState)));
    // one slot = xyz_depth[16 * nq16] then type[nq16], nq16 = the query count rounded up to 16: the slot size (17 * nq16) is
    // then a multiple of 16, so slots of several ranks can sit back to back (gathered block) with every float4 aligned
    ctx->b_results_nq = levins_result_slot_bytes(n_queries) / 17;
    CU(ctx->b_results.ensure(ctx->b_results_nq * 17));
    CU(ctx->b_flags.ensure(std::max<size_t>(n_queries, 1) * 4));
    CU(ctx->b_nn.ensure(std::max<size_t>(n_queries, 1) * 80));
    CU(ctx->b_nn_meta.ensure(std::max<size_t>(n_queries, 1) * 8));
    if (!borrowed) {
        CU(cudaMemcpyAsync(ctx->b_ins.p, ins, n_ins * sizeof(InsSample), cudaMemcpyHostToDevice, ctx->stream));
        if (n_queries) CU(cudaMemcpyAsync(ctx->b_queries.p, query_xyz, n_queries * 12, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaMemcpyAsync(ctx->b_kfs.p, kfs.data(), n_kf * sizeof(KfState), cudaMemcpyHostToDevice, ctx->stream));
    ctx->b_descs.assign(descs, descs + n_kf);
    ctx->b_nq = n_queries;
    ctx->b_npts = n_points;
    ctx->b_nins = n_ins;
    return LEVINS_OK;
}

extern "C" int levins_batch_upload(levins_ctx* ctx, const levins_point_xyzit* points, size_t n_points,
                                   const levins_ins_sample* ins, size_t n_ins, const float* query_xyz, size_t n_queries,
                                   const levins_keyframe_desc* descs, size_t n_kf) {
    if (!ctx || !points || !ins || !descs || (!query_xyz && n_queries)) return fail(ctx, LEVINS_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    // the previous batch's copies read caller memory and b_kfs_host: let them finish before either is touched again
    CU(cudaStreamSynchronize(ctx->stream));
    int r = batch_prepare(ctx, n_points, ins, n_ins, query_xyz, n_queries, descs, n_kf);
    if (r) return r;
    CU(cudaMemcpyAsync(ctx->b_pts.p, points, n_points * sizeof(PointXYZIT), cudaMemcpyHostToDevice, ctx->stream));
    ctx->b_valid = true;
    return LEVINS_OK;
}

extern "C" int levins_batch_attach_device(levins_ctx* ctx, const levins_point_xyzit* d_points, size_t n_points,
                                          const levins_ins_sample* d_ins, size_t n_ins, const float* d_query_xyz, size_t n_queries,
                                          const levins_keyframe_desc* descs, size_t n_kf) {
    if (!ctx || !d_points || !d_ins || !descs || (!d_query_xyz && n_queries)) return fail(ctx, LEVINS_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, d_points) != cudaSuccess || at.type != cudaMemoryTypeDevice || at.device != ctx->device) {
        cudaGetLastError();
        return fail(ctx, LEVINS_ERR_ARG, "d_points is not device memory of this context's GPU");
    }
    // the kernels read the pools with 16-byte vector loads and bulk copies: refuse misaligned pools instead of faulting
    if (reinterpret_cast<uintptr_t>(d_points) % 32 || reinterpret_cast<uintptr_t>(d_ins) % 8 || reinterpret_cast<uintptr_t>(d_query_xyz) % 4)
        return fail(ctx, LEVINS_ERR_ARG, "device pools must be aligned: points 32 bytes, INS samples 8, queries 4");
    // the library's own pools are released: the batch now lives in the caller's memory
    ctx->b_pts.release();
    ctx->b_ins.release();
    ctx->b_queries.release();
    int r = batch_prepare(ctx, n_points, nullptr, n_ins, nullptr, n_queries, descs, n_kf, /*borrowed=*/true);
    if (r) return r;
    ctx->b_pts_view = d_points;
    ctx->b_ins_view = d_ins;
    ctx->b_queries_view = d_query_xyz;
    ctx->b_valid = true;
    return LEVINS_OK;
}

extern "C" int levins_batch_upload_packed(levins_ctx* ctx, const levins_packed_point* points, size_t n_points,
                                          const levins_scan_desc* scans, size_t n_scans, const levins_ins_sample* ins, size_t n_ins,
                                          const float* query_xyz, size_t n_queries, const levins_keyframe_desc* descs, size_t n_kf) {
    if (!ctx || !points || !scans || !n_scans || !ins || !descs || (!query_xyz && n_queries)) return fail(ctx, LEVINS_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (n_scans > 0xffffffffull) return fail(ctx, LEVINS_ERR_ARG, "too many scans");
    unsigned long long next = 0;
    for (size_t i = 0; i < n_scans; i++) { // the scans tile the pool in order
        if (scans[i].first != next) return fail(ctx, LEVINS_ERR_ARG, "scan descriptors must tile the point pool contiguously");
        next += scans[i].count;
    }
    if (next != n_points) return fail(ctx, LEVINS_ERR_ARG, "scan descriptors do not cover the point pool");
    int r = batch_prepare(ctx, n_points, ins, n_ins, query_xyz, n_queries, descs, n_kf);
    if (r) return r;
    static_assert(sizeof(levins_packed_point) == sizeof(PackedPoint) && sizeof(levins_scan_desc) == sizeof(ScanDesc), "packed layouts");
    CU(ctx->b_packed.ensure(std::max<size_t>(n_points, 1) * sizeof(PackedPoint)));
    CU(ctx->b_scans.ensure(n_scans * sizeof(ScanDesc)));
    CU(cudaMemcpyAsync(ctx->b_scans.p, scans, n_scans * sizeof(ScanDesc), cudaMemcpyHostToDevice, ctx->stream));
    // a few pieces, each unpacked as soon as it has arrived, so that the expansion hides behind the next piece's copy when the
    // caller gave the context a stream of its own per direction; on one stream it simply follows
    const size_t piece = std::max<size_t>((n_points + 7) / 8, 1 << 20);
    for (size_t o = 0; o < n_points; o += piece) {
        const size_t n = std::min(piece, n_points - o);
        CU(cudaMemcpyAsync(ctx->b_packed.as<PackedPoint>() + o, points + o, n * sizeof(PackedPoint), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (n_points) {
        k_unpack<<<(unsigned) ((n_points + 255) / 256), 256, 0, ctx->stream>>>(ctx->b_packed.as<PackedPoint>(), n_points,
                                                                               ctx->b_scans.as<ScanDesc>(), (unsigned) n_scans,
                                                                               ctx->b_pts.as<PointXYZIT>());
        ctx->launches++;
        CU(cudaGetLastError());
    }
    ctx->b_valid = true;
    return LEVINS_OK;
}

static cudaEvent_t get_event(levins_ctx* ctx, size_t i) {
    while (ctx->ev_pool.size() <= i) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        ctx->ev_pool.push_back(ev);
    }
    return ctx->ev_pool[i];
}

extern "C" int levins_batch_run(levins_ctx* ctx) {
    if (!ctx) return LEVINS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (!ctx->b_valid) return fail(ctx, LEVINS_ERR_STATE, "levins_batch_upload has not been called");
    if (!ctx->have_ext) return fail(ctx, LEVINS_ERR_STATE, "levins_set_extrinsics has not been called");
    const size_t n_kf = ctx->b_descs.size();
    const int wave = (int) std::min<size_t>(ctx->b_wave, n_kf);
    if (ctx->b_wave_alloc < wave) {
        CU(ctx->b_zbuf.ensure((size_t) wave * CELLS * 8));
        ctx->b_gen = 255; // fresh memory: force a clear at the first generation
        ctx->b_wave_alloc = wave;
    }
    if (!ctx->side) {
        CU(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->ev_binned, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_counted, cudaEventDisableTiming));
    }
    size_t ev = 0;
    ctx->ev_spans.clear();
    auto mark = [&](int kind, bool begin) {
        if (!ctx->profiling) return;
        cudaEvent_t e = get_event(ctx, ev);
        cudaEventRecord(e, ctx->stream);
        if (begin)
            ctx->ev_spans.push_back({kind, ev, 0});
        else
            ctx->ev_spans.back().e1 = ev;
        ev++;
    };

    PrepareArgs pa{};
    pa.kfs = ctx->b_kfs.as<KfState>();
    pa.pts = ctx->batch_pts();
    pa.pts_cap = ~0ull;
    pa.ins = ctx->batch_ins();
    pa.table = ctx->b_table.as<Interval>();
    pa.ring = nullptr;
    pa.ext = ctx->ext;
    pa.c = ctx->consts;

    for (size_t k0 = 0; k0 < n_kf; k0 += wave) {
        const int nw = (int) std::min<size_t>(wave, n_kf - k0);
        uint64_t max_pts = 0, max_q = 0;
        for (int i = 0; i < nw; i++) {
            max_pts = std::max<uint64_t>(max_pts, ctx->b_descs[k0 + i].pt_count);
            max_q = std::max<uint64_t>(max_q, ctx->b_descs[k0 + i].q_count);
        }
        // The slots are about to be reused (and, once every 254 generations, cleared): the previous wave's k_count_cells on
        // the side stream must be done with them first, whether or not this wave launches a deskew kernel.
        if (k0) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_counted, 0));
        // every wave is a new z-buffer generation: the slots are reused without being cleared
        int r = next_gen(ctx, ctx->b_zbuf.as<unsigned long long>(), (size_t) ctx->b_wave_alloc * CELLS, &ctx->b_gen);
        if (r) return r;
        DeskewArgs da{};
        da.kfs = ctx->b_kfs.as<KfState>() + k0;
        da.pts = ctx->batch_pts();
        da.pts_cap = ~0ull;
        da.table = ctx->b_table.as<Interval>();
        da.zbuf = ctx->b_zbuf.as<unsigned long long>();
        da.cloud_out = nullptr;
        da.gen = ctx->b_gen;
        da.ext = ctx->ext;
        da.c = ctx->consts;
        // window search, interval tables and chunk descriptors of the wave (one block per keyframe)
        r = size_chunks(ctx, da, max_pts, (unsigned) nw);
        if (r) return r;
        pa.kfs = ctx->b_kfs.as<KfState>() + k0;
        pa.chunks = da.chunks;
        pa.chunks_per_kf = da.chunks_per_kf;
        mark(0, true);
        k_prepare<<<prepare_grid(ctx, (unsigned) nw, da.chunks_per_kf), 256, 0, ctx->stream>>>(pa);
        ctx->launches++;
        CU(cudaGetLastError());
        mark(0, false);
        if (max_pts) {
            mark(1, true);
            r = launch_deskew(ctx, da, false);
            mark(1, false);
            if (r) return r;
        }
        // M (non-empty cells per keyframe) streams the z-buffers once more; do it on the side stream while the
        // latency-bound k_associate occupies the main one
        CU(cudaEventRecord(ctx->ev_binned, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->side, ctx->ev_binned, 0));
        k_count_cells<<<dim3(32, nw), 256, 0, ctx->side>>>(ctx->b_kfs.as<KfState>() + k0, ctx->b_zbuf.as<unsigned long long>(), ctx->b_gen);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev_counted, ctx->side));
        if (max_q) {
            AssocArgs aa{};
            aa.kfs = ctx->b_kfs.as<KfState>() + k0;
            aa.queries = ctx->batch_queries();
            aa.cells = nullptr;
            aa.c2c = nullptr;
            aa.zbuf = ctx->b_zbuf.as<unsigned long long>();
            aa.pts = ctx->batch_pts();
            aa.pts_cap = ~0ull;
            aa.table = ctx->b_table.as<Interval>();
            aa.gen = ctx->b_gen;
            aa.count_cells = 0; // M is counted by k_count_cells on the side stream, concurrently with this kernel
            aa.ext = ctx->ext;
            aa.nn = ctx->b_nn.as<float4>();
            aa.nn_meta = ctx->b_nn_meta.as<uint2>();
            aa.knn_idx = nullptr;
            aa.knn_d2 = nullptr;
            aa.c = ctx->consts;
            mark(3, true);
            k_associate<true><<<dim3((unsigned) ((max_q + ASSOC_WARPS - 1) / ASSOC_WARPS), nw), ASSOC_WARPS * 32, 0, ctx->stream>>>(aa);
            mark(3, false);
            ctx->launches++;
            CU(cudaGetLastError());
        }
    }
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_counted, 0)); // join the side stream
    if (ctx->b_nq) { // plane fits for every feature of the batch in one launch (one thread per feature)
        PlaneArgs pl{};
        pl.queries = ctx->batch_queries();
        pl.nn = ctx->b_nn.as<float4>();
        pl.nn_meta = ctx->b_nn_meta.as<uint2>();
        pl.q_begin = 0;
        pl.q_count = ctx->b_nq;
        pl.type = ctx->b_results.as<int8_t>() + ctx->b_results_nq * 16;
        pl.xyz_depth = ctx->b_results.as<float4>();
        pl.flags = ctx->b_flags.as<int32_t>();
        pl.c = ctx->consts;
        if (ctx->peer_on) {
            if (ctx->b_nq > ctx->g_nq) return fail(ctx, LEVINS_ERR_CAPACITY, "more queries than levins_peer_export reserved per rank");
            if (ctx->g_slot % 16) return fail(ctx, LEVINS_ERR_STATE, "gathered slots are not 16-byte aligned");
            ctx->peer_step++;
            ctx->g_half = (int) (ctx->peer_step & 1);
            pl.peer = ctx->peer;
            pl.peer.parity_off = (unsigned long long) ctx->g_half * ctx->g_ranks * ctx->g_slot;
        }
        mark(2, true);
        k_plane<<<(unsigned) ((ctx->b_nq + 127) / 128), 128, 0, ctx->stream>>>(pl);
        mark(2, false);
        ctx->launches++;
        CU(cudaGetLastError());
        if (ctx->peer_on) {
            PeerFlags pf = ctx->pflags;
            pf.step = ctx->peer_step;
            k_peer_signal<<<1, 32, 0, ctx->stream>>>(pf);
            ctx->launches++;
            CU(cudaGetLastError());
        }
    }
    return LEVINS_OK;
}

extern "C" int levins_batch_result_pointers(levins_ctx* ctx, void** type, void** xyz_depth, void** flags) {
    if (!ctx) return LEVINS_ERR_ARG;
    if (!ctx->b_valid) return fail(ctx, LEVINS_ERR_STATE, "no batch");
    if (type) *type = ctx->b_results.as<int8_t>() + ctx->b_results_nq * 16;
    if (xyz_depth) *xyz_depth = ctx->b_results.p;
    if (flags) *flags = ctx->b_flags.p;
    return LEVINS_OK;
}

extern "C" int levins_batch_kernel_ms(levins_ctx* ctx, float ms[4], uint32_t launches[4]) {
    if (!ctx) return LEVINS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 4; i++) {
        ctx->prof_ms[i] = 0;
        ctx->prof_launches[i] = 0;
    }
    for (const auto& s : ctx->ev_spans) {
        float t = 0;
        CU(cudaEventElapsedTime(&t, ctx->ev_pool[s.e0], ctx->ev_pool[s.e1]));
        ctx->prof_ms[s.kind] += t;
        ctx->prof_launches[s.kind]++;
    }
    if (ms) memcpy(ms, ctx->prof_ms, sizeof(ctx->prof_ms));
    if (launches) memcpy(launches, ctx->prof_launches, sizeof(ctx->prof_launches));
    return LEVINS_OK;
}

extern "C" int levins_batch_download(levins_ctx* ctx, int8_t* type, float* xyz_depth, int32_t* flags, uint32_t* n_projected,
                                     uint32_t* m) {
    if (!ctx) return LEVINS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    if (!ctx->b_valid) return fail(ctx, LEVINS_ERR_STATE, "no batch");
    const size_t nq = ctx->b_nq, n_kf = ctx->b_descs.size();
    if (type && nq)
        CU(cudaMemcpyAsync(type, ctx->b_results.as<int8_t>() + ctx->b_results_nq * 16, nq, cudaMemcpyDeviceToHost, ctx->stream));
    if (xyz_depth && nq) CU(cudaMemcpyAsync(xyz_depth, ctx->b_results.p, nq * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (flags && nq) CU(cudaMemcpyAsync(flags, ctx->b_flags.p, nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<KfState> kfs;
    if (n_projected || m) {
        kfs.resize(n_kf);
        CU(cudaMemcpyAsync(kfs.data(), ctx->b_kfs.p, n_kf * sizeof(KfState), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < kfs.size(); k++) {
        if (n_projected) n_projected[k] = kfs[k].n_proj;
        if (m) m[k] = kfs[k].m;
    }
    return update_cull_policy(ctx);
}

#include "exchange_host.inc"
