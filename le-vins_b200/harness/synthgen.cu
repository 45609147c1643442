// synthgen.cu -- HARNESS, not product: device-side generator of the BASELINE config-5 workload ("10 000 independent
// keyframes of shape 3, generated on device from per-keyframe seeds to avoid 10 k x 84 MB of H2D").  Built as its own library
// (libsynthgen_b200.so) and used by bench.py and tests only; liblevins_b200.so consumes its output through
// levins_batch_attach_device.
//
// Every value is produced with integer hashing and IEEE +, -, *, /, sqrt only (no contraction, no transcendental function at
// run time: the scan directions come from a table the host computes once), so levins/synthgen.py reproduces the same bits
// with numpy and the CPU oracle can be run on exactly the keyframes the GPU processed.
//
// Keyframe k (seed s): an Ouster OS1-128-like scanner (128 rings x 1024 columns, one timestamp per column) rides a vehicle
// that moves with constant velocity and turns with a rational-quaternion yaw; the world is an axis-aligned box corridor in
// the LiDAR frame at the first scan (ground, two side walls, front and back walls) whose dimensions come from the seed.
#include <cstdint>
#include <cuda_runtime.h>

namespace {

struct PointXYZIT {
    float x, y, z, pad0;
    float intensity, pad1;
    double time;
};
struct InsSample {
    double time, state_time;
    double p[3];
    double q[4];
};
static_assert(sizeof(PointXYZIT) == 32 && sizeof(InsSample) == 72, "ABI layouts");

__host__ __device__ inline uint32_t mix(uint32_t h) { // lowbias32
    h ^= h >> 16;
    h *= 0x7feb352du;
    h ^= h >> 15;
    h *= 0x846ca68bu;
    h ^= h >> 16;
    return h;
}
__host__ __device__ inline uint32_t hash3(uint32_t a, uint32_t b, uint32_t c) { return mix(mix(mix(a) ^ b) ^ c); }
// uniform in [0, 1) with 24 bits, exactly representable in f32
__host__ __device__ inline float u01(uint32_t h) { return (float) (h >> 8) * (1.0f / 16777216.0f); }

struct GenArgs {
    uint32_t seed, first_kf, n_kf, scans_per_kf, n_ins, n_feat;
    double t_base;       // GPS second of week of keyframe 0's first scan
    double kf_period;    // seconds between consecutive keyframes' windows
    double td_b_l;
    const float* dirs;   // [1024][128][3] unit directions, column-major in time
    PointXYZIT* points;  // [n_kf][scans_per_kf * 131072]
    InsSample* ins;      // [n_kf][n_ins]
    float* queries;      // [n_kf][n_feat][3]
    double intr[4];      // fx fy cx cy
    uint32_t width, height;
};

constexpr uint32_t RINGS = 128, COLS = 1024, PER_SCAN = RINGS * COLS;
constexpr double SCAN_PERIOD = 0.1, INS_RATE_DT = 0.005;

__global__ void __launch_bounds__(256) k_gen_points(const GenArgs a) {
    const uint32_t kf = blockIdx.y, gk = a.first_kf + kf;
    const uint32_t per_kf = a.scans_per_kf * PER_SCAN;
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= per_kf) return;
    const uint32_t scan = i / PER_SCAN, in_scan = i - scan * PER_SCAN, col = in_scan / RINGS;
    // corridor of this keyframe
    const float half_w = 4.0f + 8.0f * u01(hash3(a.seed, gk, 1));
    const float front = 15.0f + 60.0f * u01(hash3(a.seed, gk, 2));
    const float back = 10.0f + 40.0f * u01(hash3(a.seed, gk, 3));
    const float ground = 1.5f + 0.8f * u01(hash3(a.seed, gk, 4));
    const float roof = 4.0f + 20.0f * u01(hash3(a.seed, gk, 5));
    const float dx = a.dirs[3 * in_scan], dy = a.dirs[3 * in_scan + 1], dz = a.dirs[3 * in_scan + 2];
    float r = 250.0f; // nothing hit: beyond the far range gate
    if (dz < -1e-3f) r = fminf(r, ground / -dz);
    if (dz > 1e-3f) r = fminf(r, roof / dz);
    if (dy > 1e-3f) r = fminf(r, half_w / dy);
    if (dy < -1e-3f) r = fminf(r, half_w / -dy);
    if (dx > 1e-3f) r = fminf(r, front / dx);
    if (dx < -1e-3f) r = fminf(r, back / -dx);
    const uint32_t h = hash3(a.seed ^ 0x9e3779b9u, gk, i);
    r = r + (u01(h) * 0.02f - 0.01f);               // +-1 cm of range noise
    if ((h & 1023u) == 7u) r = 0.3f + 0.5f * u01(mix(h)); // a sprinkle inside the near range gate
    PointXYZIT p;
    p.x = dx * r;
    p.y = dy * r;
    p.z = dz * r;
    p.pad0 = 0.f;
    p.intensity = (float) (h & 255u);
    p.pad1 = 0.f;
    const double t0k = a.t_base + (double) gk * a.kf_period;
    p.time = (t0k + (double) scan * SCAN_PERIOD) + (double) col * (SCAN_PERIOD / (double) COLS);
    a.points[(size_t) kf * per_kf + i] = p;
}

__global__ void __launch_bounds__(256) k_gen_ins_queries(const GenArgs a) {
    const uint32_t kf = blockIdx.y, gk = a.first_kf + kf;
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    const double t0k = a.t_base + (double) gk * a.kf_period;
    if (i < a.n_ins) {
        // the INS window starts 0.6 s before the first point's INS time (t + td_b_l)
        const double t = ((t0k + a.td_b_l) - 0.6) + (double) i * INS_RATE_DT;
        const double tau = t - t0k;
        const double vx = 2.0 + 12.0 * (double) u01(hash3(a.seed, gk, 6));
        const double vy = -0.5 + 1.0 * (double) u01(hash3(a.seed, gk, 7));
        const double w = -0.15 + 0.3 * (double) u01(hash3(a.seed, gk, 8)); // half yaw rate: q ~ (0, 0, w tau, 1)
        const double u = w * tau;
        const double n = sqrt(1.0 + u * u);
        InsSample s;
        s.time = t;
        s.state_time = t;
        s.p[0] = vx * tau;
        s.p[1] = vy * tau;
        s.p[2] = 0.02 * tau;
        s.q[0] = 0.0;
        s.q[1] = 0.0;
        s.q[2] = u / n;
        s.q[3] = 1.0 / n;
        a.ins[(size_t) kf * a.n_ins + i] = s;
    }
    if (i < a.n_feat) { // tracking.cc:1103-1106: uniform pixels with a 5-px border; Camera::pixel2unitcam; cast to float
        const uint32_t hx = hash3(a.seed ^ 0x51ed270bu, gk, 2 * i), hy = hash3(a.seed ^ 0x51ed270bu, gk, 2 * i + 1);
        const float px = 5.0f + (float) (hx % (a.width - 10u)), py = 5.0f + (float) (hy % (a.height - 10u));
        const double y = ((double) py - a.intr[3]) / a.intr[1];
        const double x = ((double) px - a.intr[2]) / a.intr[0];
        const double s = sqrt((x * x + y * y) + 1.0);
        float* q = a.queries + ((size_t) kf * a.n_feat + i) * 3;
        q[0] = (float) (x / s);
        q[1] = (float) (y / s);
        q[2] = (float) (1.0 / s);
    }
}

} // namespace

// Generates keyframes [first_kf, first_kf + n_kf) into caller-owned DEVICE buffers on `stream`.  Returns 0 or a cudaError_t.
extern "C" int synthgen_ouster_batch(int device, void* stream, uint32_t seed, uint32_t first_kf, uint32_t n_kf, uint32_t scans_per_kf,
                                     uint32_t n_ins, uint32_t n_feat, double t_base, double kf_period, double td_b_l,
                                     const float* d_dirs, void* d_points, void* d_ins, float* d_queries, const double intr[4],
                                     uint32_t width, uint32_t height) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int) e;
    GenArgs a{};
    a.seed = seed;
    a.first_kf = first_kf;
    a.n_kf = n_kf;
    a.scans_per_kf = scans_per_kf;
    a.n_ins = n_ins;
    a.n_feat = n_feat;
    a.t_base = t_base;
    a.kf_period = kf_period;
    a.td_b_l = td_b_l;
    a.dirs = d_dirs;
    a.points = static_cast<PointXYZIT*>(d_points);
    a.ins = static_cast<InsSample*>(d_ins);
    a.queries = d_queries;
    for (int i = 0; i < 4; i++) a.intr[i] = intr[i];
    a.width = width;
    a.height = height;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t per_kf = scans_per_kf * PER_SCAN;
    k_gen_points<<<dim3((per_kf + 255) / 256, n_kf), 256, 0, st>>>(a);
    const uint32_t m = n_ins > n_feat ? n_ins : n_feat;
    k_gen_ins_queries<<<dim3((m + 255) / 256, n_kf), 256, 0, st>>>(a);
    return (int) cudaGetLastError();
}
