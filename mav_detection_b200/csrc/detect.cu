// Derotation, Focus-of-Expansion RANSAC, radial residual / masks / metrics and connected components
// for sm_100a.  Follows /root/reference/src/detector.py:70-117, focus_of_expansion.py:32-86,150-184,
// processor.py:306-362, im_helpers.py:55-84,244-252, utils.py:183-197.
//
// Parity-critical float64 arithmetic uses the explicit round-to-nearest intrinsics (__dmul_rn,
// __dadd_rn, ...) so that nvcc cannot contract a*b+c into an FMA: NumPy evaluates every ufunc with
// one rounding per operation and the masks must be bit-exact.
#include <math.h>

#include "common.cuh"

namespace mavd {

// ------------------------------------------------------------------------------------------------
// BGR -> gray: cv2.cvtColor(COLOR_BGR2GRAY) on uint8 == (B*3735 + G*19235 + R*9798 + 16384) >> 15
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bgr2gray_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray,
                                                      int64_t n) {
    int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if (i4 + 4 <= n && ((uintptr_t)bgr & 3) == 0 && ((uintptr_t)gray & 3) == 0) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(bgr + i4 * 3);
        uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        uint8_t v[12];
        v[0] = a; v[1] = a >> 8; v[2] = a >> 16; v[3] = a >> 24;
        v[4] = b; v[5] = b >> 8; v[6] = b >> 16; v[7] = b >> 24;
        v[8] = c; v[9] = c >> 8; v[10] = c >> 16; v[11] = c >> 24;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t g = (v[3 * k] * 3735u + v[3 * k + 1] * 19235u + v[3 * k + 2] * 9798u + 16384u) >> 15;
            out |= g << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(gray + i4) = out;
    } else {
        for (int64_t i = i4; i < n && i < i4 + 4; ++i)
            gray[i] = (uint8_t)((bgr[3 * i] * 3735u + bgr[3 * i + 1] * 19235u + bgr[3 * i + 2] * 9798u + 16384u) >> 15);
    }
}

int bgr2gray_run(const uint8_t* d_bgr, uint8_t* d_gray, int64_t n, cudaStream_t s) {
    int64_t threads = (n + 3) / 4;
    bgr2gray_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(d_bgr, d_gray, n);
    MAVD_LAUNCHED();
    return MAVD_OK;
}

// ------------------------------------------------------------------------------------------------
// Derotation (detector.py:88-101), evaluated with NumPy's operation order, one rounding per op.
// ------------------------------------------------------------------------------------------------
struct Derot {
    double o0, o1, o2;  // omega = ang / dt
    double s0, s1;      // w*dt/2, h*dt/2
    double w, h;
    int on;
};

__device__ __forceinline__ Derot make_derot(const mavd_imu& imu, int w, int h) {
    Derot d;
    d.on = imu.derotate;
    d.o0 = __ddiv_rn(imu.ang[0], imu.dt);
    d.o1 = __ddiv_rn(imu.ang[1], imu.dt);
    d.o2 = __ddiv_rn(imu.ang[2], imu.dt);
    d.s0 = __ddiv_rn(__dmul_rn((double)w, imu.dt), 2.0);
    d.s1 = __ddiv_rn(__dmul_rn((double)h, imu.dt), 2.0);
    d.w = (double)w;
    d.h = (double)h;
    return d;
}

__device__ __forceinline__ void derot_at(const Derot& d, int x, int y, double& r0, double& r1) {
    const double xn = __dmul_rn(-__dsub_rn(__ddiv_rn((double)x, d.w), 0.5), 2.0);
    const double yn = __dmul_rn(-__dsub_rn(__ddiv_rn((double)y, d.h), 0.5), 2.0);
    // +o0*xn*yn - o1*xn**2 - o1 + o2*yn
    double t = __dmul_rn(__dmul_rn(d.o0, xn), yn);
    t = __dsub_rn(t, __dmul_rn(d.o1, __dmul_rn(xn, xn)));
    t = __dsub_rn(t, d.o1);
    t = __dadd_rn(t, __dmul_rn(d.o2, yn));
    r0 = __dmul_rn(t, d.s0);
    // -o2*xn + o0 + o0*yn**2 - o1*xn*yn
    double u = __dmul_rn(-d.o2, xn);
    u = __dadd_rn(u, d.o0);
    u = __dadd_rn(u, __dmul_rn(d.o0, __dmul_rn(yn, yn)));
    u = __dsub_rn(u, __dmul_rn(__dmul_rn(d.o1, xn), yn));
    r1 = __dmul_rn(u, d.s1);
}

__global__ void __launch_bounds__(256) derotate_kernel(const float2* __restrict__ flow, const mavd_imu* __restrict__ imu,
                                                      int w, int h, double2* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= w) return;
    const Derot d = make_derot(imu[f], w, h);
    const size_t o = ((size_t)f * h + y) * w + x;
    const float2 v = flow[o];
    double r0 = 0.0, r1 = 0.0;
    if (d.on) derot_at(d, x, y, r0, r1);
    out[o] = make_double2(d.on ? __dsub_rn((double)v.x, r0) : (double)v.x,
                          d.on ? __dsub_rn((double)v.y, r1) : (double)v.y);
}

int derotate_run(mavd_handle H, const float* d_flow, int n, const mavd_imu* d_imu, double* d_out, cudaStream_t s) {
    dim3 g(ceil_div(H->cfg.width, 256), H->cfg.height, n);
    derotate_kernel<<<g, 256, 0, s>>>((const float2*)d_flow, d_imu, H->cfg.width, H->cfg.height, (double2*)d_out);
    MAVD_LAUNCHED();
    return MAVD_OK;
}

// ------------------------------------------------------------------------------------------------
// FoE: 1000 flow-line intersections + consensus scoring.  One 1024-thread CTA per frame.
//   phase A  thread i builds intersection i (float64, utils.line_intersection order of operations)
//   phase B  order-preserving compaction of rows with x != 0 (ballot + warp offsets)
//   phase C  thread i counts estimates within ransac_threshold of estimate i
//   phase D  arg-max with first-maximum tie-break (warp shuffles, then across warps)
// sqrt(s) < T is evaluated as s < s*, s* = the smallest double whose correctly rounded square root
// is >= T (computed on the host), which is exactly equivalent and saves 10^6 double square roots.
// ------------------------------------------------------------------------------------------------
constexpr int kNP = MAVD_N_SAMPLE_PAIRS;

__global__ void __launch_bounds__(1024) foe_kernel(const float2* __restrict__ flow, const mavd_imu* __restrict__ imu,
                                                  const int32_t* __restrict__ samples, int w, int h,
                                                  double mag_thr, double sq_thr, double* __restrict__ foe,
                                                  int32_t* __restrict__ ninter) {
    __shared__ double2 E[kNP];
    __shared__ int warp_cnt[32];
    __shared__ unsigned long long warp_best[32];
    const int f = blockIdx.x, i = threadIdx.x, lane = i & 31, wid = i >> 5;
    const Derot d = make_derot(imu[f], w, h);
    const float2* fl = flow + (size_t)f * w * h;
    const int32_t* sm = samples + (size_t)f * MAVD_SAMPLES_PER_FRAME;

    bool valid = false;
    double ex = 0.0, ey = 0.0;
    if (i < kNP) {
        const int y1 = sm[i], y2 = sm[i + kNP], x1 = sm[2 * kNP + i], x2 = sm[3 * kNP + i];
        const float2 a = fl[(size_t)y1 * w + x1], b = fl[(size_t)y2 * w + x2];
        double f1x, f1y, f2x, f2y;
        bool keep;
        if (d.on) {
            double r0, r1;
            derot_at(d, x1, y1, r0, r1);
            f1x = __dsub_rn((double)a.x, r0); f1y = __dsub_rn((double)a.y, r1);
            derot_at(d, x2, y2, r0, r1);
            f2x = __dsub_rn((double)b.x, r0); f2y = __dsub_rn((double)b.y, r1);
            const double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(f2x, f2x), __dmul_rn(f2y, f2y)));
            keep = !(mag < mag_thr);
        } else {
            f1x = a.x; f1y = a.y; f2x = b.x; f2y = b.y;
            const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(b.x, b.x), __fmul_rn(b.y, b.y)));
            keep = !(mag < (float)mag_thr);
        }
        if (keep) {
            const double cx1 = x1, cy1 = y1, cx2 = x2, cy2 = y2;
            const double p1x = __dadd_rn(f1x, cx1), p1y = __dadd_rn(f1y, cy1);
            const double p2x = __dadd_rn(f2x, cx2), p2y = __dadd_rn(f2y, cy2);
            const double xd0 = __dsub_rn(cx1, p1x), xd1 = __dsub_rn(cx2, p2x);
            const double yd0 = __dsub_rn(cy1, p1y), yd1 = __dsub_rn(cy2, p2y);
            const double div = __dsub_rn(__dmul_rn(xd0, yd1), __dmul_rn(xd1, yd0));
            if (div != 0.0) {
                const double d0 = __dsub_rn(__dmul_rn(cx1, p1y), __dmul_rn(cy1, p1x));
                const double d1 = __dsub_rn(__dmul_rn(cx2, p2y), __dmul_rn(cy2, p2x));
                ex = __ddiv_rn(__dsub_rn(__dmul_rn(d0, xd1), __dmul_rn(d1, xd0)), div);
                ey = __ddiv_rn(__dsub_rn(__dmul_rn(d0, yd1), __dmul_rn(d1, yd0)), div);
                valid = (ex != 0.0);  // also true for NaN, as in NumPy
            }
        }
    }
    // ---- phase B ----
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    if (wid == 0) {
        int v = warp_cnt[lane], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        warp_cnt[lane] = incl - v;
        if (lane == 31) warp_best[0] = (unsigned long long)incl;  // total, parked here for a moment
    }
    __syncthreads();
    const int K = (int)warp_best[0];
    if (valid) E[warp_cnt[wid] + __popc(bal & ((1u << lane) - 1u))] = make_double2(ex, ey);
    __syncthreads();
    // ---- phase C ----
    int score = -1;
    if (i < K) {
        const double2 me = E[i];
        int cnt = 0;
        for (int j = 0; j < K; ++j) {
            const double2 o = E[j];
            const double dx = __dsub_rn(o.x, me.x), dy = __dsub_rn(o.y, me.y);
            const double s2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            cnt += (s2 < sq_thr) ? 1 : 0;
        }
        score = cnt - 1;
    }
    // ---- phase D: maximise (score, -index) ----
    unsigned long long key = (score > 0) ? (((unsigned long long)(unsigned)score << 32) | (unsigned)(0x7fffffff - i)) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o);
        key = t > key ? t : key;
    }
    __syncthreads();
    if (lane == 0) warp_best[wid] = key;
    __syncthreads();
    if (wid == 0) {
        key = warp_best[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long t = __shfl_xor_sync(0xffffffffu, key, o);
            key = t > key ? t : key;
        }
        if (lane == 0) {
            double bx = 0.0, by = 0.0;
            if (key != 0ull) {
                const int best = 0x7fffffff - (int)(key & 0xffffffffu);
                bx = E[best].x;
                by = E[best].y;
            }
            foe[2 * f] = bx;
            foe[2 * f + 1] = by;
            ninter[f] = K;
        }
    }
}

// smallest double s with correctly-rounded sqrt(s) >= T  (so that  sqrt(s) < T  <=>  s < s*)
static double sqrt_threshold(double T) {
    if (!(T > 0.0)) return 0.0;
    double s = T * T;
    while (sqrt(s) >= T) s = nextafter(s, 0.0);
    while (sqrt(s) < T) s = nextafter(s, INFINITY);
    return s;
}

int foe_run(mavd_handle H, const float* d_flow, int n, const mavd_imu* d_imu, const mavd_detect_params& prm,
            const int32_t* d_samples, double* d_foe, int32_t* d_ninter, cudaStream_t s) {
    ProfScope ps(&H->prof, MAVD_PROF_FOE, s);
    foe_kernel<<<n, 1024, 0, s>>>((const float2*)d_flow, d_imu, d_samples, H->cfg.width, H->cfg.height,
                                  prm.magnitude_threshold, sqrt_threshold(prm.ransac_threshold), d_foe, d_ninter);
    MAVD_LAUNCHED();
    return MAVD_OK;
}

// ------------------------------------------------------------------------------------------------
// Residual angle phi + the two masks + the integer reductions behind FrameResult.
// ------------------------------------------------------------------------------------------------
struct ResidualPrm {
    double dyn_offset, dyn_base, dyn_gain, dyn_min_mag, fixed_min_mag, fixed_angle;
};

__device__ __forceinline__ unsigned long long dmax_key(double v) { return (unsigned long long)__double_as_longlong(v); }

__global__ void __launch_bounds__(256) seg_max_kernel(const uint8_t* __restrict__ seg, int64_t seg_stride, int64_t npx,
                                                     int* __restrict__ seg_max) {
    const int f = blockIdx.y;
    const uint8_t* p = seg + (size_t)f * seg_stride;
    int mx = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (int64_t)gridDim.x * blockDim.x)
        mx = max(mx, (int)p[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(seg_max + f, mx);
}

struct BlockStats {
    long long n_total, n_fixed, pos, neg, tp_t, fp_t, tp_f, fp_f;
    int x0, y0, x1, y1;
    double sfx, sfy, maxphi;
};

__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool F64>
__global__ void __launch_bounds__(256) residual_kernel(const float2* __restrict__ flow, const mavd_imu* __restrict__ imu,
                                                      const double* __restrict__ foe, int w, int h, ResidualPrm prm,
                                                      const uint8_t* __restrict__ sky, int64_t sky_stride,
                                                      const uint8_t* __restrict__ seg, int64_t seg_stride,
                                                      const int* __restrict__ seg_max, void* __restrict__ phi_out,
                                                      uint8_t* __restrict__ total_out, uint8_t* __restrict__ fixed_out,
                                                      char* __restrict__ stats_base, size_t stats_stride) {
    const int f = blockIdx.z;
    const mavd_imu im = imu[f];
    // both precisions are launched for every batch; each handles only its own frames
    if ((im.derotate != 0) != F64) return;
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    const bool in = (x < w && y < h);
    const size_t o = in ? ((size_t)y * w + x) : 0;
    const size_t fo = (size_t)f * w * h + o;
    bool m_total = false, m_fixed = false;
    double phi_d = 0.0, fdx = 0.0, fdy = 0.0;
    if (in) {
        const float2 v = flow[fo];
        const bool not_sky = sky ? (sky[(size_t)f * sky_stride + o] == 0) : true;
        const double foex = foe[2 * f], foey = foe[2 * f + 1];
        if (F64) {
            const Derot d = make_derot(im, w, h);
            double r0, r1;
            derot_at(d, x, y, r0, r1);
            fdx = __dsub_rn((double)v.x, r0);
            fdy = __dsub_rn((double)v.y, r1);
            const double d2x = __dsub_rn((double)x, foex), d2y = __dsub_rn((double)y, foey);
            const double a = __dsqrt_rn(__dadd_rn(__dmul_rn(fdx, fdx), __dmul_rn(fdy, fdy)));
            const double b = __dsqrt_rn(__dadd_rn(__dmul_rn(d2x, d2x), __dmul_rn(d2y, d2y)));
            const double ab = __dmul_rn(a, b);
            const double norm = (ab != ab) ? ab : fmax(1e-6, ab);
            double c = __ddiv_rn(__dadd_rn(__dmul_rn(fdx, d2x), __dmul_rn(fdy, d2y)), norm);
            if (c == c) c = fmin(fmax(c, -1.0), 1.0);
            double ang = acos(c);
            if (ang != ang) ang = 0.0;
            const double phi = __dmul_rn(ang, 180.0 / 3.141592653589793238462643383279502884);
            phi_d = phi;
            const double t = __dadd_rn(prm.dyn_base, __ddiv_rn(prm.dyn_gain, a));
            const bool amax = phi > __dadd_rn(prm.dyn_offset, t);
            const bool amin = phi < __dsub_rn(prm.dyn_offset, t);
            m_total = (a > prm.dyn_min_mag) && not_sky && (amin || amax);
            m_fixed = (phi > prm.fixed_angle) && (a > prm.fixed_min_mag) && not_sky;
            if (phi_out) reinterpret_cast<double*>(phi_out)[fo] = phi;
        } else {
            // frame_index < 1: the flow stays float32 and so does every NumPy temporary
            const float fx = v.x, fy = v.y;
            fdx = fx; fdy = fy;
            const float d2x = (float)__dsub_rn((double)x, foex), d2y = (float)__dsub_rn((double)y, foey);
            const float a = __fsqrt_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
            const float b = __fsqrt_rn(__fadd_rn(__fmul_rn(d2x, d2x), __fmul_rn(d2y, d2y)));
            const float ab = __fmul_rn(a, b);
            const float norm = (ab != ab) ? ab : fmaxf(1e-6f, ab);
            float c = __fdiv_rn(__fadd_rn(__fmul_rn(fx, d2x), __fmul_rn(fy, d2y)), norm);
            if (c == c) c = fminf(fmaxf(c, -1.f), 1.f);
            float ang = (float)acos((double)c);
            if (ang != ang) ang = 0.f;
            const float phi = __fmul_rn(ang, 180.0f / 3.141592653589793238462643383279502884f);
            phi_d = phi;
            const float t = __fadd_rn((float)prm.dyn_base, __fdiv_rn((float)prm.dyn_gain, a));
            const bool amax = phi > __fadd_rn((float)prm.dyn_offset, t);
            const bool amin = phi < __fsub_rn((float)prm.dyn_offset, t);
            m_total = (a > (float)prm.dyn_min_mag) && not_sky && (amin || amax);
            m_fixed = (phi > (float)prm.fixed_angle) && (a > (float)prm.fixed_min_mag) && not_sky;
            if (phi_out) reinterpret_cast<float*>(reinterpret_cast<double*>(phi_out) + (size_t)f * w * h)[o] = phi;
        }
        if (total_out) total_out[fo] = m_total ? 1 : 0;
        if (fixed_out) fixed_out[fo] = m_fixed ? 1 : 0;
    }
    if (!stats_base) return;

    // ---- block reduction of the integer statistics ----
    long long c_tot = m_total, c_fix = m_fixed, c_pos = 0, c_neg = 0, c_tpt = 0, c_fpt = 0, c_tpf = 0, c_fpf = 0;
    int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -1, by1 = -1;
    double sfx = 0.0, sfy = 0.0;
    if (in && seg) {
        const int g = seg[(size_t)f * seg_stride + o];
        c_pos = g > 127;
        c_neg = (255 - g) > 127;
        c_tpt = m_total && g >= 1;
        c_fpt = m_total && g <= 254;
        c_tpf = m_fixed && g >= 1;
        c_fpf = m_fixed && g <= 254;
        const double thr = 0.1 * (double)seg_max[f];
        if ((double)g > thr) { bx0 = bx1 = x; by0 = by1 = y; }
        if (g > 127) { sfx = fdx; sfy = fdy; }
    }
    __shared__ BlockStats red[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    c_tot = warp_sum(c_tot); c_fix = warp_sum(c_fix);
    unsigned long long mk = dmax_key(phi_d);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, mk, s);
        mk = t > mk ? t : mk;
    }
    if (seg) {
        c_pos = warp_sum(c_pos); c_neg = warp_sum(c_neg); c_tpt = warp_sum(c_tpt); c_fpt = warp_sum(c_fpt);
        c_tpf = warp_sum(c_tpf); c_fpf = warp_sum(c_fpf);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            bx0 = min(bx0, __shfl_xor_sync(0xffffffffu, bx0, s));
            by0 = min(by0, __shfl_xor_sync(0xffffffffu, by0, s));
            bx1 = max(bx1, __shfl_xor_sync(0xffffffffu, bx1, s));
            by1 = max(by1, __shfl_xor_sync(0xffffffffu, by1, s));
            sfx += __shfl_xor_sync(0xffffffffu, sfx, s);
            sfy += __shfl_xor_sync(0xffffffffu, sfy, s);
        }
    }
    if (lane == 0) {
        BlockStats& b = red[wid];
        b.n_total = c_tot; b.n_fixed = c_fix; b.pos = c_pos; b.neg = c_neg; b.tp_t = c_tpt; b.fp_t = c_fpt;
        b.tp_f = c_tpf; b.fp_f = c_fpf; b.x0 = bx0; b.y0 = by0; b.x1 = bx1; b.y1 = by1; b.sfx = sfx; b.sfy = sfy;
        b.maxphi = __longlong_as_double((long long)mk);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        BlockStats t = red[0];
        for (int k = 1; k < 8; ++k) {
            const BlockStats& b = red[k];
            t.n_total += b.n_total; t.n_fixed += b.n_fixed; t.pos += b.pos; t.neg += b.neg; t.tp_t += b.tp_t;
            t.fp_t += b.fp_t; t.tp_f += b.tp_f; t.fp_f += b.fp_f;
            t.x0 = min(t.x0, b.x0); t.y0 = min(t.y0, b.y0); t.x1 = max(t.x1, b.x1); t.y1 = max(t.y1, b.y1);
            t.sfx += b.sfx; t.sfy += b.sfy; t.maxphi = fmax(t.maxphi, b.maxphi);
        }
        mavd_frame_stats* st = reinterpret_cast<mavd_frame_stats*>(stats_base + (size_t)f * stats_stride);
        typedef unsigned long long ull;
        if (t.n_total) atomicAdd((ull*)&st->n_total, (ull)t.n_total);
        if (t.n_fixed) atomicAdd((ull*)&st->n_fixed, (ull)t.n_fixed);
        if (t.maxphi > 0.0) atomicMax((ull*)&st->max_phi, dmax_key(t.maxphi));
        if (seg) {
            if (t.pos) atomicAdd((ull*)&st->positives, (ull)t.pos);
            if (t.neg) atomicAdd((ull*)&st->negatives, (ull)t.neg);
            if (t.tp_t) atomicAdd((ull*)&st->tp_total, (ull)t.tp_t);
            if (t.fp_t) atomicAdd((ull*)&st->fp_total, (ull)t.fp_t);
            if (t.tp_f) atomicAdd((ull*)&st->tp_fixed, (ull)t.tp_f);
            if (t.fp_f) atomicAdd((ull*)&st->fp_fixed, (ull)t.fp_f);
            if (t.x1 >= 0) {
                // seg_bbox was initialised to {INT_MAX, INT_MAX, -1, -1} by stats_init_kernel
                atomicMin(&st->seg_bbox[0], t.x0); atomicMin(&st->seg_bbox[1], t.y0);
                atomicMax(&st->seg_bbox[2], t.x1); atomicMax(&st->seg_bbox[3], t.y1);
            }
            if (t.sfx != 0.0) atomicAdd(&st->seg_flow_sum[0], t.sfx);
            if (t.sfy != 0.0) atomicAdd(&st->seg_flow_sum[1], t.sfy);
        }
    }
}

__global__ void stats_init_kernel(char* stats_base, size_t stats_stride, int n, int* seg_max) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    mavd_frame_stats* st = reinterpret_cast<mavd_frame_stats*>(stats_base + (size_t)f * stats_stride);
    mavd_frame_stats z;
    memset(&z, 0, sizeof(z));
    z.seg_bbox[0] = z.seg_bbox[1] = 0x7fffffff;
    z.seg_bbox[2] = z.seg_bbox[3] = -1;
    *st = z;
    if (seg_max) seg_max[f] = 0;
}

__global__ void stats_final_kernel(char* stats_base, size_t stats_stride, int n) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    mavd_frame_stats* st = reinterpret_cast<mavd_frame_stats*>(stats_base + (size_t)f * stats_stride);
    if (st->seg_bbox[2] < 0) st->seg_bbox[0] = st->seg_bbox[1] = st->seg_bbox[2] = st->seg_bbox[3] = -1;
}

int residual_run(mavd_handle H, const float* d_flow, int n, const mavd_imu* d_imu, const mavd_detect_params& p,
                 const double* d_foe, const uint8_t* d_sky, int64_t sky_stride, const uint8_t* d_seg,
                 int64_t seg_stride, void* d_phi, uint8_t* d_total, uint8_t* d_fixed, mavd_frame_stats* d_stats,
                 size_t stats_stride, int run_f64, int run_f32, cudaStream_t s) {
    ProfScope ps(&H->prof, MAVD_PROF_RESIDUAL, s);
    const int w = H->cfg.width, h = H->cfg.height;
    int* seg_max = reinterpret_cast<int*>(H->d_scan);  // scratch: n ints
    if (d_stats) {
        stats_init_kernel<<<ceil_div(n, 128), 128, 0, s>>>((char*)d_stats, stats_stride, n, d_seg ? seg_max : nullptr);
        MAVD_LAUNCHED();
        if (d_seg) {
            dim3 g(148 * 2, n);
            seg_max_kernel<<<g, 256, 0, s>>>(d_seg, seg_stride, (int64_t)w * h, seg_max);
            MAVD_LAUNCHED();
        }
    }
    ResidualPrm rp{p.dyn_offset, p.dyn_base, p.dyn_gain, p.dyn_min_mag, p.fixed_min_mag, p.fixed_angle};
    dim3 g(ceil_div(w, 64), ceil_div(h, 4), n);
    if (run_f64) {
        residual_kernel<true><<<g, 256, 0, s>>>((const float2*)d_flow, d_imu, d_foe, w, h, rp, d_sky, sky_stride, d_seg,
                                                seg_stride, seg_max, d_phi, d_total, d_fixed, (char*)d_stats, stats_stride);
        MAVD_LAUNCHED();
    }
    if (run_f32) {
        residual_kernel<false><<<g, 256, 0, s>>>((const float2*)d_flow, d_imu, d_foe, w, h, rp, d_sky, sky_stride, d_seg,
                                                 seg_stride, seg_max, d_phi, d_total, d_fixed, (char*)d_stats, stats_stride);
        MAVD_LAUNCHED();
    }
    if (d_stats && d_seg) {
        stats_final_kernel<<<ceil_div(n, 128), 128, 0, s>>>((char*)d_stats, stats_stride, n);
        MAVD_LAUNCHED();
    }
    return MAVD_OK;
}

// ------------------------------------------------------------------------------------------------
// Connected components (8-connectivity), union-find with the smaller raster index as the root, so a
// component's root is its first pixel in raster order and ranking the roots gives canonical labels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int* parent, int i) {
    int p = parent[i];
    while (p != i) {
        i = p;
        p = parent[i];
    }
    return i;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;
    }
}

__global__ void __launch_bounds__(256) ccl_init_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent, int npx) {
    const size_t base = (size_t)blockIdx.y * npx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x)
        parent[base + i] = mask[base + i] ? i : -1;
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent, int w, int h) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= w || y >= h) return;
    const size_t base = (size_t)blockIdx.z * w * h;
    const uint8_t* m = mask + base;
    int* par = parent + base;
    const int i = y * w + x;
    if (!m[i]) return;
    if (x > 0 && m[i - 1]) uf_union(par, i, i - 1);
    if (y > 0) {
        if (m[i - w]) uf_union(par, i, i - w);
        else {
            // with the pixel above set, both diagonals are already joined through it
            if (x > 0 && m[i - w - 1]) uf_union(par, i, i - w - 1);
            if (x + 1 < w && m[i - w + 1]) uf_union(par, i, i - w + 1);
        }
    }
}

constexpr int CCL_CHUNK = 4096;  // pixels per block in the root-ranking passes

__global__ void __launch_bounds__(256) ccl_flatten_count_kernel(int* __restrict__ parent, int npx, int* __restrict__ chunk_cnt,
                                                               int n_chunks) {
    const size_t base = (size_t)blockIdx.y * npx;
    int* par = parent + base;
    const int c0 = blockIdx.x * CCL_CHUNK;
    int cnt = 0;
    for (int i = c0 + threadIdx.x; i < min(c0 + CCL_CHUNK, npx); i += 256) {
        int p = par[i];
        if (p >= 0) {
            int r = uf_find(par, i);
            par[i] = r;   // benign race: every writer stores a valid ancestor, roots never change here
            cnt += (r == i);
        }
    }
    __shared__ int wsum[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += wsum[k];
        chunk_cnt[(size_t)blockIdx.y * n_chunks + blockIdx.x] = t;
    }
}

// exclusive scan of the per-chunk root counts, one block per frame; writes the total label count
__global__ void __launch_bounds__(1024) ccl_scan_kernel(int* __restrict__ chunk_cnt, int n_chunks, char* nlabels_base,
                                                       size_t nlabels_stride) {
    __shared__ int wtot[32];
    __shared__ int carry_s;
    int* c = chunk_cnt + (size_t)blockIdx.x * n_chunks;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_chunks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n_chunks ? c[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wtot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int wv = wtot[lane], wi = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wtot[lane] = wi - wv;
        }
        __syncthreads();
        const int carry = carry_s;
        if (i < n_chunks) c[i] = carry + wtot[wid] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wtot[wid] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *reinterpret_cast<int32_t*>(nlabels_base + (size_t)blockIdx.x * nlabels_stride) = carry_s;
}

// rank[root] = canonical label of the component rooted at `root`
__global__ void __launch_bounds__(256) ccl_rank_kernel(const int* __restrict__ parent, int npx, const int* __restrict__ chunk_off,
                                                      int n_chunks, int* __restrict__ rank) {
    const size_t base = (size_t)blockIdx.y * npx;
    const int* par = parent + base;
    int* rk = rank + base;
    const int c0 = blockIdx.x * CCL_CHUNK;
    __shared__ int wsum[8];
    __shared__ int running;
    if (threadIdx.x == 0) running = chunk_off[(size_t)blockIdx.y * n_chunks + blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int s0 = c0; s0 < min(c0 + CCL_CHUNK, npx); s0 += 256) {
        const int i = s0 + threadIdx.x;
        const bool root = (i < npx) && (par[i] == i);
        const unsigned bal = __ballot_sync(0xffffffffu, root);
        if (lane == 0) wsum[wid] = __popc(bal);
        __syncthreads();
        int off = running;
        for (int k = 0; k < wid; ++k) off += wsum[k];
        if (root) rk[i] = off + __popc(bal & ((1u << lane) - 1u)) + 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < 8; ++k) t += wsum[k];
            running += t;
        }
        __syncthreads();
    }
}

__global__ void ccl_boxes_init_kernel(int32_t* boxes, size_t boxes_stride, int max_boxes, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * max_boxes) return;
    int f = i / max_boxes, b = i - f * max_boxes;
    int32_t* p = boxes + (size_t)f * boxes_stride + b * 5;
    p[0] = 0x7fffffff; p[1] = 0x7fffffff; p[2] = -1; p[3] = -1; p[4] = 0;
}

__global__ void __launch_bounds__(256) ccl_relabel_kernel(int* __restrict__ labels, const int* __restrict__ rank, int w, int h,
                                                         int32_t* __restrict__ boxes, size_t boxes_stride, int max_boxes) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= w || y >= h) return;
    const size_t base = (size_t)blockIdx.z * w * h;
    const int i = y * w + x;
    const int r = labels[base + i];
    int lab = 0;
    if (r >= 0) {
        lab = rank[base + r];
        if (boxes && lab <= max_boxes) {
            int32_t* b = boxes + (size_t)blockIdx.z * boxes_stride + (lab - 1) * 5;
            atomicMin(b + 0, x); atomicMin(b + 1, y); atomicMax(b + 2, x); atomicMax(b + 3, y); atomicAdd(b + 4, 1);
        }
    }
    labels[base + i] = lab;
}

__global__ void ccl_boxes_final_kernel(int32_t* boxes, size_t boxes_stride, int max_boxes, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * max_boxes) return;
    int f = i / max_boxes, b = i - f * max_boxes;
    int32_t* p = boxes + (size_t)f * boxes_stride + b * 5;
    if (p[4] == 0) { p[0] = p[1] = p[2] = p[3] = 0; }
    else { p[2] = p[2] - p[0] + 1; p[3] = p[3] - p[1] + 1; }
}

int ccl_run(mavd_handle H, const uint8_t* d_mask, int n, int32_t* d_labels, int32_t* d_boxes, size_t boxes_stride,
            int max_boxes, int32_t* d_n_labels, size_t nlabels_stride, cudaStream_t s) {
    ProfScope ps(&H->prof, MAVD_PROF_CCL, s);
    const int w = H->cfg.width, h = H->cfg.height, npx = w * h;
    const int n_chunks = ceil_div(npx, CCL_CHUNK);
    int* rank = H->d_scan;                                  // [n][npx]
    int* chunk_cnt = H->d_scan + (size_t)H->cfg.max_pairs * npx;  // [n][n_chunks]
    ccl_init_kernel<<<dim3(148 * 4, n), 256, 0, s>>>(d_mask, d_labels, npx);
    MAVD_LAUNCHED();
    dim3 g(ceil_div(w, 64), ceil_div(h, 4), n);
    ccl_merge_kernel<<<g, 256, 0, s>>>(d_mask, d_labels, w, h);
    MAVD_LAUNCHED();
    ccl_flatten_count_kernel<<<dim3(n_chunks, n), 256, 0, s>>>(d_labels, npx, chunk_cnt, n_chunks);
    MAVD_LAUNCHED();
    ccl_scan_kernel<<<n, 1024, 0, s>>>(chunk_cnt, n_chunks, (char*)d_n_labels, nlabels_stride);
    MAVD_LAUNCHED();
    ccl_rank_kernel<<<dim3(n_chunks, n), 256, 0, s>>>(d_labels, npx, chunk_cnt, n_chunks, rank);
    MAVD_LAUNCHED();
    if (d_boxes) {
        ccl_boxes_init_kernel<<<ceil_div(n * max_boxes, 128), 128, 0, s>>>(d_boxes, boxes_stride, max_boxes, n);
        MAVD_LAUNCHED();
    }
    ccl_relabel_kernel<<<g, 256, 0, s>>>(d_labels, rank, w, h, d_boxes, boxes_stride, max_boxes);
    MAVD_LAUNCHED();
    if (d_boxes) {
        ccl_boxes_final_kernel<<<ceil_div(n * max_boxes, 128), 128, 0, s>>>(d_boxes, boxes_stride, max_boxes, n);
        MAVD_LAUNCHED();
    }
    return MAVD_OK;
}

}  // namespace mavd
