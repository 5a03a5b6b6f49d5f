// rtw_device.cuh — device functions of the render hot path (sm_100a).
//
// One section per reference module; every function cites the reference lines it restates
// (paths relative to /root/reference).  Arithmetic is f32 except where the reference's f64 is needed for
// conditioning: the sphere discriminant (sphere_hit) runs in f64 — B200 executes FP64 at half FP32 rate.
#ifndef RTW_DEVICE_CUH
#define RTW_DEVICE_CUH

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "rtw_types.h"

namespace rtwd {

#define RTW_DEV __device__ __forceinline__

// Scene features a kernel instantiation must support.  The render kernel is compiled in 8 variants (launch_all picks
// the smallest superset of what the scene uses): the all-features kernel is 64 KB of SASS, the one for (moving) spheres
// with solid / checker textures (book-1: config 1) 34 KB, and the instruction cache (L1.5 = 32 KB per SM) is the first
// thing a megakernel runs out of (ncu: stall_no_instruction).
// FEAT_XFORM: some primitive sits under Translate / RotateY (hit records replay the wrapper chain); FEAT_RXFORM: some RECT
// does, so rect tests must move the ray into object space (final_scene: only its baked spheres are instanced).
enum { FEAT_RECT = 1, FEAT_XFORM = 2, FEAT_MEDIA = 4, FEAT_NOISE = 8, FEAT_IMAGE = 16, FEAT_RXFORM = 32, FEAT_ALL = 63 };
#define RTW_PI_F 3.14159265358979323846f

// ------------------------------------------------------------------------------------------------
// src/math.rs — Vector3 (:12-20) and its operators (:147-266) in f32
// ------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
RTW_DEV V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
RTW_DEV V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
RTW_DEV V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
RTW_DEV V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
RTW_DEV V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
RTW_DEV V3 operator*(float s, V3 a) { return mk(a.x * s, a.y * s, a.z * s); }
RTW_DEV V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
RTW_DEV float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }       // :82-84
RTW_DEV float length_squared(V3 a) { return dot(a, a); }                                    // :86-88
// MUFU-backed sqrt / reciprocal (max relative error 2^-23, PTX ISA) for values that only SELECT a candidate: the IEEE
// sequences behind sqrtf() and 1.0f/x cost ~10 instructions each plus an out-of-line slow path (85 SASS instructions
// for the one sqrtf of sphere_root).  Accepted sphere hits are re-derived in f64 (finalize_hit), box tests are padded.
RTW_DEV float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
RTW_DEV float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
RTW_DEV V3 normalize(V3 v) { float inv = rsqrtf(length_squared(v)); return inv * v; }        // :102-104 ((1/len)*v, :260-266); MUFU.RSQ, 2 ulp
RTW_DEV V3 reflect(V3 v, V3 n) { return v - (2.0f * dot(v, n)) * n; }                       // :106-108
RTW_DEV V3 refract(V3 uv, V3 n, float etai_over_etat) {                                      // :110-117
    float cos_theta = fminf(dot(-uv, n), 1.0f);
    V3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    V3 r_out_parallel = (-sqrtf(fabsf(1.0f - length_squared(r_out_perp)))) * n;
    return r_out_perp + r_out_parallel;
}
RTW_DEV bool near_zero(V3 v) { const float S = 1e-8f; return fabsf(v.x) < S && fabsf(v.y) < S && fabsf(v.z) < S; }  // :134-137
RTW_DEV void sphere_uv(V3 p, float& u, float& v) {                                           // :288-300
    // theta = acos(-p.y) evaluated as atan2(|(x,z)|, -y): identical for a unit vector, but keeps full relative
    // precision at the poles where acos of an f32 within 1 ulp of +-1 loses everything.
    // The two atan2f share one inlined copy (the loop is kept rolled on purpose: code size).
    float ang[2];
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        const float a = k ? -p.z : sqrtf(p.x * p.x + p.z * p.z), b = k ? p.x : -p.y;
        const float r = atan2f(a, b);
        if (k) ang[1] = r; else ang[0] = r;
    }
    u = (ang[1] + RTW_PI_F) * (1.0f / (2.0f * RTW_PI_F));
    v = ang[0] * (1.0f / RTW_PI_F);
}
// sin for bounded arguments: reduction by pi/2 + the usual f32 minimax pair, WITHOUT the Payne-Hanek path sinf carries
// for |x| > 1e5 (150 instructions).  Used for the marble phase scale*z + 10*turb, a function of scene coordinates.
// |err| < 1.5e-7 up to |x| ~ 1e9.
// The phase arrives in f64 and is reduced in f64 (one DFMA against pi/2 and its tail): at |x| ~ 200 an f32 phase is
// already quantised to 1.5e-5 — the whole 1e-5 budget of the texture value.
RTW_DEV float sin_bounded(double x) {
    const double kd = rint(x * 0.63661977236758134);
    const int q = __double2int_rn(kd);
    const float r = (float)fma(kd, -6.123233995736766e-17, fma(kd, -1.5707963267948966, x));
    const float s = r * r;
    const float sn = fmaf(fmaf(fmaf(-1.95152959e-4f, s, 8.33216087e-3f), s, -1.66666546e-1f), s * r, r);
    const float cs = fmaf(fmaf(fmaf(fmaf(2.44331571e-5f, s, -1.38873163e-3f), s, 4.16666457e-2f), s, -0.5f), s, 1.0f);
    const float v = (q & 1) ? cs : sn;
    return (q & 2) ? -v : v;
}

// ------------------------------------------------------------------------------------------------
// RNG — replaces rand::thread_rng (src/math.rs:268-280).  Philox4x32-10, key = seed,
// counter = (draw block, bounce, pixel, sample); a draw is (word >> 8) * 2^-24 (identical in the oracle).
// ------------------------------------------------------------------------------------------------
RTW_DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                           uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// The key is the same for every block of a launch: its 10-round schedule travels in the kernel parameters
// (DParams::philox_rk, filled per launch from the seed by make_params; kernel parameters live in the constant bank) and
// feeds the LOP3s directly — 20 integer adds fewer per block than bumping the key.  Nothing is shared between
// launches, so renders of different scenes / seeds never serialise on a symbol.
RTW_DEV void philox4x32_10_rk(const uint32_t* __restrict__ rk, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                              uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ rk[2 * r]; c1 = lo1; c2 = hi0 ^ c3 ^ rk[2 * r + 1]; c3 = lo0;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// Stream layout (the CPU checker of the test suite uses the same one): scalar draws of a bounce are words of blocks 0, 1, 2, ... of its
// counter line; the two rejection loops take their attempts from blocks of their own, indexed by the ATTEMPT: unit sphere
// (src/math.rs:51-58) attempt a = words 0..2 of block 0x40000000 + a, unit disk (:69-76) attempt a = words 2(a & 1),
// 2(a & 1) + 1 of block 0x20000000 + (a >> 1).  They do not advance the scalar stream.  Any lane can therefore evaluate any
// attempt of any other lane from (pixel, sample, bounce) alone — coop_unit_sphere below.
RTW_DEV float u01(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
RTW_DEV void philox_sphere_attempt(const uint32_t* __restrict__ rk, uint32_t a, uint32_t bounce, uint32_t pixel, uint32_t sample, float& x, float& y, float& z) {
    uint32_t w0, w1, w2, w3;
    philox4x32_10_rk(rk, 0x40000000u + a, bounce, pixel, sample, w0, w1, w2, w3);
    x = u01(w0); y = u01(w1); z = u01(w2);
}

struct PhiloxRng {
    uint32_t pixel, sample, bounce, draw;
    uint32_t w0, w1, w2, w3;
    const uint32_t* rk;                            // the launch's key schedule: &prm.philox_rk[0] of a __grid_constant__ DParams
    RTW_DEV void bind(const DParams& prm) { rk = prm.philox_rk; }
    RTW_DEV void start(uint32_t px, uint32_t s) { pixel = px; sample = s; bounce = 0; draw = 0; }      // after bind()
    RTW_DEV void init(const DParams& prm, uint32_t px, uint32_t s) { bind(prm); start(px, s); }
    RTW_DEV void set_bounce(uint32_t b) { bounce = b; draw = 0; }
    RTW_DEV float next() {                               // random_double()
        uint32_t i = draw & 3u;
        if (i == 0) philox4x32_10_rk(rk, draw >> 2, bounce, pixel, sample, w0, w1, w2, w3);
        ++draw;
        uint32_t w = i == 0 ? w0 : (i == 1 ? w1 : (i == 2 ? w2 : w3));
        return (float)(w >> 8) * (1.0f / 16777216.0f);
    }
    // Two / three CONSECUTIVE draws behind one inlined copy of the block function (every next() call site carries its
    // own 70-instruction copy; the kernel is instruction-cache bound).  Same words as next() called 2 / 3 times.
    // `single`: only the first draw is consumed (camera time, the last draw of bounce 0).
    RTW_DEV void next2(float& x, float& y, bool single = false) {
        const uint32_t i = draw & 3u;
        const uint32_t ox = i == 1 ? w1 : (i == 2 ? w2 : w3), oy = i == 1 ? w2 : w3;
        if (i == 0 || (i == 3 && !single)) philox4x32_10_rk(rk, (draw + 1) >> 2, bounce, pixel, sample, w0, w1, w2, w3);
        const uint32_t wx = i == 0 ? w0 : ox, wy = i == 0 ? w1 : (i == 3 ? w0 : oy);
        draw += single ? 1u : 2u;
        x = (float)(wx >> 8) * (1.0f / 16777216.0f); y = (float)(wy >> 8) * (1.0f / 16777216.0f);
    }
    RTW_DEV void sphere_attempt(uint32_t a, float& x, float& y, float& z) const { philox_sphere_attempt(rk, a, bounce, pixel, sample, x, y, z); }
    // two unit-disk attempts (2 blk, 2 blk + 1) from one block; `second` = the second pair is a real attempt
    RTW_DEV void disk_pair(uint32_t blk, float& x0, float& y0, float& x1, float& y1, bool& second) const {
        uint32_t a0, a1, a2, a3;
        philox4x32_10_rk(rk, 0x20000000u + blk, bounce, pixel, sample, a0, a1, a2, a3);
        x0 = u01(a0); y0 = u01(a1); x1 = u01(a2); y1 = u01(a3); second = true;
    }
    RTW_DEV void next3(float& x, float& y, float& z) {
        const uint32_t i = draw & 3u;
        const uint32_t ox = i == 1 ? w1 : (i == 2 ? w2 : w3), oy = i == 1 ? w2 : w3, oz = w3;
        if (i != 1) philox4x32_10_rk(rk, (draw + 3) >> 2, bounce, pixel, sample, w0, w1, w2, w3);
        const uint32_t wx = i == 0 ? w0 : ox, wy = i == 0 ? w1 : (i == 3 ? w0 : oy), wz = i == 0 ? w2 : (i == 1 ? oz : (i == 2 ? w0 : w1));
        draw += 3;
        x = (float)(wx >> 8) * (1.0f / 16777216.0f); y = (float)(wy >> 8) * (1.0f / 16777216.0f); z = (float)(wz >> 8) * (1.0f / 16777216.0f);
    }
};

// explicit stream (parity hooks): `stride` f64 draws per item, consumed in order
struct StreamRng {
    const double* xi; int n; int draw;
    RTW_DEV void set_bounce(uint32_t) {}
    RTW_DEV void sphere_attempt(uint32_t, float& x, float& y, float& z) { x = next(); y = next(); z = next(); }     // sequential, like the reference
    RTW_DEV void disk_pair(uint32_t, float& x0, float& y0, float& x1, float& y1, bool& second) { x0 = next(); y0 = next(); x1 = 0.f; y1 = 0.f; second = false; }
    RTW_DEV float next() { float v = draw < n ? (float)xi[draw] : 0.5f; ++draw; return v; }
    RTW_DEV void next2(float& x, float& y, bool single = false) { x = next(); y = single ? 0.5f : next(); }
    RTW_DEV void next3(float& x, float& y, float& z) { x = next(); y = next(); z = next(); }
};

template <class R> RTW_DEV float rng_range(R& g, float a, float b) { return a + (b - a) * g.next(); }   // :273-276
template <class R> RTW_DEV V3 random_in_unit_sphere(R& g) {                                               // :51-58, draw order x,y,z :43-49
#pragma unroll 1
    for (uint32_t a = 0;; ++a) {
        float x, y, z;
        g.sphere_attempt(a, x, y, z);
        V3 p = mk(-1.0f + 2.0f * x, -1.0f + 2.0f * y, -1.0f + 2.0f * z);                  // random_range(-1, 1) :273-276
        if (length_squared(p) < 1.0f) return p;
    }
}

// The same sample, drawn by the whole WARP for the lanes that `need` one (all 32 lanes must call).  Every lane tries its
// own attempt 0; then, round by round, the lanes that are done evaluate the NEXT attempts of the lanes that are not (an
// attempt is a Philox block addressed by (pixel, sample, bounce, attempt): any lane can compute it).  With f failed lanes
// each gets 32 / f attempts per round, so the loop ends after ~2 rounds instead of running until the unluckiest lane's
// private loop does (acceptance pi/6: the slowest of 16 lanes needs ~5 attempts).  Same attempts in the same order as
// random_in_unit_sphere: the first accepted attempt wins, bit-identical to the scalar loop (and to the oracle).
// `scr`: 32 ints of per-warp shared memory.
__constant__ uint32_t c_stride_mask[33] = {
    0u, 0xffffffffu, 0x55555555u, 0x09249249u, 0x11111111u, 0x02108421u, 0x01041041u, 0x00204081u, 0x01010101u, 0x00040201u, 0x00100401u,
    0x00000801u, 0x00001001u, 0x00002001u, 0x00004001u, 0x00008001u, 0x00010001u,
    1u, 1u, 1u, 1u, 1u, 1u, 1u, 1u, 1u, 1u, 1u, 1u, 1u, 1u, 1u, 1u};
RTW_DEV V3 coop_unit_sphere(bool need, const PhiloxRng& g, int lane, int* scr) {
    V3 p = mk(0.f, 0.f, 0.f);
    bool ok = !need;
    if (need) {
        float x, y, z;
        g.sphere_attempt(0u, x, y, z);
        p = mk(-1.0f + 2.0f * x, -1.0f + 2.0f * y, -1.0f + 2.0f * z);
        ok = length_squared(p) < 1.0f;
    }
    unsigned failed = __ballot_sync(0xffffffffu, !ok);
    uint32_t base = 1u;
#pragma unroll 1
    while (failed) {
        const int nf = __popc(failed);
        const int idx = __popc(failed & ((1u << lane) - 1u));            // my index among the failed lanes (if I am one)
        if (!ok) scr[idx] = lane;
        __syncwarp();
        const int j = __float2int_rz(((float)lane + 0.5f) * rcp_approx((float)nf));      // lane / nf
        const int t = lane - j * nf;                                                     // lane % nf
        const int per = __float2int_rz(32.5f * rcp_approx((float)nf));                    // attempts per failed lane this round
        const int tgt = scr[t];
        const uint32_t tp = __shfl_sync(0xffffffffu, g.pixel, tgt), ts = __shfl_sync(0xffffffffu, g.sample, tgt), tb = __shfl_sync(0xffffffffu, g.bounce, tgt);
        float x, y, z;
        philox_sphere_attempt(g.rk, base + (uint32_t)j, tb, tp, ts, x, y, z);
        const V3 q = mk(-1.0f + 2.0f * x, -1.0f + 2.0f * y, -1.0f + 2.0f * z);
        const bool okq = j < per && length_squared(q) < 1.0f;
        const unsigned won = __ballot_sync(0xffffffffu, okq);
        // the workers of failed lane #idx are lanes idx, idx + nf, idx + 2 nf, ...: the lowest one that accepted holds the
        // earliest accepted attempt
        const unsigned mine = won & (c_stride_mask[nf] << idx);
        const bool got = !ok && mine != 0u;
        const int src = got ? __ffs(mine) - 1 : lane;
        const float qx = __shfl_sync(0xffffffffu, q.x, src), qy = __shfl_sync(0xffffffffu, q.y, src), qz = __shfl_sync(0xffffffffu, q.z, src);
        if (got) { p = mk(qx, qy, qz); ok = true; }
        base += (uint32_t)per;
        failed = __ballot_sync(0xffffffffu, !ok);
    }
    return p;
}
template <class R> RTW_DEV V3 random_unit_vector(R& g) { return normalize(random_in_unit_sphere(g)); }   // :78-80

// ------------------------------------------------------------------------------------------------
// src/ray.rs (:3-7, :19-21) — directions are never normalised
// ------------------------------------------------------------------------------------------------
struct Ray { V3 o, d; float time; };
RTW_DEV V3 ray_at(const Ray& r, float t) { return mk(fmaf(t, r.d.x, r.o.x), fmaf(t, r.d.y, r.o.y), fmaf(t, r.d.z, r.o.z)); }

// ------------------------------------------------------------------------------------------------
// src/camera.rs:58-66 — get_ray.  Draw order: disk loop (x, y per iteration, src/math.rs:69-76), then time.
// ------------------------------------------------------------------------------------------------
// One loop, one next2() call site: pass 0 draws the pixel jitter (src/main.rs:518-519; skipped when the caller
// supplies s, t), the next passes the lens-disk attempts, the last one the shutter time.
template <bool JITTER, class R> RTW_DEV Ray camera_ray_loop(const DCamera& c, float s, float t, float px, float py, float wm1, float hm1, R& g) {
    // scalar draws of bounce 0: pixel jitter u, v (src/main.rs:518-519), then the shutter time (camera.rs:64) — one block;
    // the lens-disk loop draws pairs of attempts from its own blocks (stream layout above)
    float tm;
    if (JITTER) {
        float a, b, d;
        g.next3(a, b, d);
        s = __fdividef(px + a, wm1); t = __fdividef(py + b, hm1);                             // 2 ulp: sub-pixel jitter
        tm = c.time0 + (c.time1 - c.time0) * d;
    }
    float rx = 0.f, ry = 0.f;
#pragma unroll 1
    for (uint32_t blk = 0;; ++blk) {
        float x0, y0, x1, y1; bool second;
        g.disk_pair(blk, x0, y0, x1, y1, second);
        rx = -1.0f + 2.0f * x0; ry = -1.0f + 2.0f * y0;
        if (rx * rx + ry * ry < 1.0f) break;
        if (second) {
            rx = -1.0f + 2.0f * x1; ry = -1.0f + 2.0f * y1;
            if (rx * rx + ry * ry < 1.0f) break;
        }
    }
    if (!JITTER) tm = c.time0 + (c.time1 - c.time0) * g.next();         // explicit stream: disk draws first, then the time
    rx *= c.lens_radius; ry *= c.lens_radius;
    V3 offset = mk(c.ux * rx + c.wx * ry, c.uy * rx + c.wy * ry, c.uz * rx + c.wz * ry);
    Ray r;
    r.o = mk(c.ox + offset.x, c.oy + offset.y, c.oz + offset.z);
    r.d = mk(c.lx + s * c.hx + t * c.vx - offset.x, c.ly + s * c.hy + t * c.vy - offset.y, c.lz + s * c.hz + t * c.vz - offset.z);
    r.time = tm;
    return r;
}
template <class R> RTW_DEV Ray camera_get_ray(const DCamera& c, float s, float t, R& g) { return camera_ray_loop<false>(c, s, t, 0.f, 0.f, 1.f, 1.f, g); }

// ------------------------------------------------------------------------------------------------
// src/perlin.rs — noise (:32-68), perlin_interp (:70-94), turb (:96-108).  Smoothstep is applied twice and the
// weight vector uses the once-smoothed fraction, exactly like the reference.
// ------------------------------------------------------------------------------------------------
RTW_DEV float perlin_noise(const uint8_t* __restrict__ tbl, V3 p) {
    const float4* ranvec = reinterpret_cast<const float4*>(tbl);
    const uint8_t* perm = tbl + 256 * 16;
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    u = u * u * (3.0f - 2.0f * u); v = v * v * (3.0f - 2.0f * v); w = w * w * (3.0f - 2.0f * w);
    int i = __float2int_rz(fx), j = __float2int_rz(fy), k = __float2int_rz(fz);   // saturating like Rust `as i32`
    float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
    float accum = 0.0f;
#pragma unroll                                      // (rolling the outer loop: 110 instructions less, 4 % slower; all 8 corners as one loop: -240,
                                                    //  Perlin scenes 75 % slower, final_scene +0.5 %: profiles/r2_ax_final_scene_code_size.log)
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int dk = 0; dk < 2; ++dk) {
                int h = __ldg(perm + ((i + di) & 255)) ^ __ldg(perm + 256 + ((j + dj) & 255)) ^ __ldg(perm + 512 + ((k + dk) & 255));
                float4 g = __ldg(ranvec + h);
                float fi = (float)di, fj = (float)dj, fk = (float)dk;
                float wt = (di ? uu : 1.0f - uu) * (dj ? vv : 1.0f - vv) * (dk ? ww : 1.0f - ww);
                accum += wt * (g.x * (u - fi) + g.y * (v - fj) + g.z * (w - fk));
            }
    return accum;
}
// The octave sum runs in f64 (7 DFMA; the octave weights and coordinate doublings are powers of two, exact either way):
// 10 * turb feeds a sine, and an f32 sum of seven f32 noise values alone leaves ~1e-6 on the phase.
RTW_DEV double perlin_turb(const uint8_t* __restrict__ tbl, V3 p, int depth) {
    double accum = 0.0; float weight = 1.0f;
    for (int i = 0; i < depth; ++i) {
        accum = fma((double)weight, (double)perlin_noise(tbl, p), accum);
        weight *= 0.5f;
        p = p * 2.0f;
    }
    return fabs(accum);
}

// ------------------------------------------------------------------------------------------------
// src/texture.rs:30-75 — get_color_value
// ------------------------------------------------------------------------------------------------
template <int F = FEAT_ALL>
RTW_DEV V3 texture_value(const DScene& sc, int tex, float u, float v, V3 p) {
    const float4* tp = reinterpret_cast<const float4*>(sc.texs + tex);
    float4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
    int kind = __float_as_int(t1.w);
    if (kind == TEX_SOLID) return mk(t0.x, t0.y, t0.z);
    if (kind == TEX_CHECKER) {                                                             // :35-42
        // Only the SIGN of sin(10x)·sin(10y)·sin(10z) is used: sin(a) < 0 iff floor(a/pi) is odd, and the product is
        // zero (-> even) only when a coordinate is 0.  Evaluated in f64 (three DMUL + F2I) so the cell boundaries sit
        // where the f64 reference puts them; three inlined sinf() were 700 SASS instructions of this kernel.
        const int odd = (__double2int_rd(10.0 * (double)p.x * 0.31830988618379067) ^ __double2int_rd(10.0 * (double)p.y * 0.31830988618379067) ^
                         __double2int_rd(10.0 * (double)p.z * 0.31830988618379067)) & 1;
        const bool zero = p.x == 0.0f || p.y == 0.0f || p.z == 0.0f;
        return (odd && !zero) ? mk(t1.x, t1.y, t1.z) : mk(t0.x, t0.y, t0.z);
    }
    if (!(F & (FEAT_NOISE | FEAT_IMAGE))) return mk(t0.x, t0.y, t0.z);
    int4 ti = __ldg(reinterpret_cast<const int4*>(tp + 2));
    if ((F & FEAT_NOISE) && (kind == TEX_NOISE || !(F & FEAT_IMAGE))) {                                                               // :43-45
        float c = 0.5f * (1.0f + sin_bounded(fma((double)t0.w, (double)p.z, 10.0 * perlin_turb(sc.perlin + (size_t)ti.x * RTW_PERLIN_BYTES, p, 7))));
        return mk(c, c, c);
    }
    // TEX_IMAGE                                                                           // :46-73
    float uu = fminf(fmaxf(u, 0.0f), 1.0f);
    float vv = 1.0f - fminf(fmaxf(v, 0.0f), 1.0f);
    int i = (int)(uu * (float)ti.y), j = (int)(vv * (float)ti.z);
    if (i >= ti.y) i = ti.y - 1;
    if (j >= ti.z) j = ti.z - 1;
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    const uint8_t* px = sc.image + (size_t)ti.x + (size_t)j * ti.w + (size_t)i * 3;
    const float s = 1.0f / 255.0f;
    return mk(s * (float)__ldg(px), s * (float)__ldg(px + 1), s * (float)__ldg(px + 2));
}

// ------------------------------------------------------------------------------------------------
// src/hittable.rs — primitive intersection
// ------------------------------------------------------------------------------------------------
struct HitRec {               // HitRecord :6-15
    V3 p, normal; float t; int front; int mat; float u, v;
};

// Per-segment ray data kept in registers during traversal.
#ifndef RTW_LAZY_F64
#define RTW_LAZY_F64 0
#endif
struct TRay {
    V3 o, d; float time;
#if RTW_LAZY_F64
    float inv_a;                           // f64 copies are re-derived at each sphere test (frees 14 registers)
    RTW_DEV double gox() const { return (double)o.x; } RTW_DEV double goy() const { return (double)o.y; } RTW_DEV double goz() const { return (double)o.z; }
    RTW_DEV double gdx() const { return (double)d.x; } RTW_DEV double gdy() const { return (double)d.y; } RTW_DEV double gdz() const { return (double)d.z; }
    RTW_DEV double ga() const { double x = d.x, y = d.y, z = d.z; return x * x + y * y + z * z; }
#else
    double ox, oy, oz, dx, dy, dz, a;      // f64 copies for the sphere discriminant; a = |d|^2
    float inv_a;                           // 1 / a
    RTW_DEV double gox() const { return ox; } RTW_DEV double goy() const { return oy; } RTW_DEV double goz() const { return oz; }
    RTW_DEV double gdx() const { return dx; } RTW_DEV double gdy() const { return dy; } RTW_DEV double gdz() const { return dz; }
    RTW_DEV double ga() const { return a; }
#endif
};
RTW_DEV TRay make_tray(const Ray& r) {
    TRay t; t.o = r.o; t.d = r.d; t.time = r.time;
#if RTW_LAZY_F64
    t.inv_a = rcp_approx((float)t.ga());
#else
    t.ox = r.o.x; t.oy = r.o.y; t.oz = r.o.z; t.dx = r.d.x; t.dy = r.d.y; t.dz = r.d.z;
    t.a = t.dx * t.dx + t.dy * t.dy + t.dz * t.dz;
    t.inv_a = rcp_approx((float)t.a);
#endif
    return t;
}

RTW_DEV void load_prim_center(const DPrim* __restrict__ pp, int type, float inv_dt, float time, double& cx, double& cy, double& cz, double& r) {
    const double2* q = reinterpret_cast<const double2*>(pp);
    double2 c01 = __ldg(q), c23 = __ldg(q + 1);
    cx = c01.x; cy = c01.y; cz = c23.x; r = c23.y;
    if (type == PRIM_MOVING_SPHERE) {                                                      // get_center_at_time :556-558
        const float4 mv = __ldg(reinterpret_cast<const float4*>(pp) + 2);                  // centre1 - centre0, time0
        double s = (double)((time - mv.w) * inv_dt);
        cx = fma(s, (double)mv.x, cx); cy = fma(s, (double)mv.y, cy); cz = fma(s, (double)mv.z, cz);
    }
}

// sphere_hit :254-288 — roots only.  oc, half_b, c and the discriminant in f64 (the reference's precision),
// then the numerically stable root pair in f32 (q = -(half_b + sign(half_b) sqrt(disc)); roots q/a and c/q).
// Branch-free: a warp processes 32 different leaves, an early-out would only add divergence.
// Returns the accepted root or NaN.
RTW_DEV float sphere_root(const DPrim* __restrict__ pp, int type, float inv_dt, const TRay& r, float t_lo, float t_hi, bool self, float* far_root = nullptr) {
    double cx, cy, cz, rad;
    load_prim_center(pp, type, inv_dt, r.time, cx, cy, cz, rad);
    double ocx = r.gox() - cx, ocy = r.goy() - cy, ocz = r.goz() - cz;
    double half_b = ocx * r.gdx() + ocy * r.gdy() + ocz * r.gdz();
    double c = ocx * ocx + ocy * ocy + ocz * ocz - rad * rad;
    double disc = half_b * half_b - r.ga() * c;
    float discf = (float)disc, hb = (float)half_b, cf = (float)c;
    float sq = sqrt_approx(fmaxf(discf, 0.0f));
    bool neg = hb < 0.0f;
    float q = neg ? sq - hb : -hb - sq;
    float tq = q * r.inv_a, tc = __fdividef(cf, q);
    float t_near = neg ? tc : tq, t_far = neg ? tq : tc;
    bool near_ok = t_near >= t_lo && t_near <= t_hi;                                         // :266-273
    float root = near_ok ? t_near : t_far;
    // `self`: the ray STARTS on this sphere (it is the primitive the path just scattered from).  In the reference's
    // f64 that root is ~1e-13 < t_min and is never returned; in f32 the stored origin is off the surface by one
    // coordinate quantum and the root can exceed t_min for grazing directions.  Exact geometry: the root at the
    // origin is c/q (c ~ 0); only the other one, q/a, can be a real hit.
    if (self) root = tq;
    if (far_root) *far_root = discf >= 0.0f ? t_far : CUDART_NAN_F;       // the other crossing (ConstantMedium boundary)
    bool ok = discf >= 0.0f && root >= t_lo && root <= t_hi;
    return ok ? root : CUDART_NAN_F;
}

RTW_DEV void xform_ray(const DScene& sc, int xf, const TRay& r, V3& o, V3& d) {            // Translate :233, RotateY :390-394 (composed)
    const DXform* x = sc.xforms + xf;
    float2 m = __ldg(reinterpret_cast<const float2*>(x));
    const double2* dp = reinterpret_cast<const double2*>(&x->d_cos);
    double2 cs = __ldg(dp), bxy = __ldg(dp + 1); double bz = __ldg(&x->d_bz);
    // origin in f64 (then rounded once): keeps (k - o_k) accurate when the origin is close to a rect's plane
    o = mk((float)(cs.x * r.gox() - cs.y * r.goz() + bxy.x), (float)(r.goy() + bxy.y), (float)(cs.y * r.gox() + cs.x * r.goz() + bz));
    d = mk(m.x * r.d.x - m.y * r.d.z, r.d.y, m.y * r.d.x + m.x * r.d.z);
}

// xy/xz/yz_rect_hit :308-384 — t only.  The ray is in the rect's object space.  The three variants differ only in
// which component plays k / a / b: selected with predicated moves so that lanes holding different rect kinds (the six
// faces of a box sit in neighbouring leaves) run ONE instruction stream.
RTW_DEV float rect_root(const DPrim* __restrict__ pp, int type, V3 o, V3 d, float t_lo, float t_hi) {
    const float4* q = reinterpret_cast<const float4*>(pp);
    float4 ab = __ldg(q); float k = __ldg(reinterpret_cast<const float*>(q + 1));
    const bool xy = type == PRIM_XY, yz = type == PRIM_YZ;
    const float ok = xy ? o.z : (yz ? o.x : o.y), dk = xy ? d.z : (yz ? d.x : d.y);
    const float oa = yz ? o.y : o.x, da = yz ? d.y : d.x;
    const float ob = xy ? o.y : o.z, db = xy ? d.y : d.z;
    float t = __fdividef(k - ok, dk);               // MUFU.RCP + FMUL, 2 ulp (IEEE division: 10 instructions + slow path per test)
    float a = oa + t * da, b = ob + t * db;
    bool okk = t >= t_lo && t <= t_hi && !(a < ab.x || a > ab.y || b < ab.z || b > ab.w);
    return okk ? t : CUDART_NAN_F;
}

// Box (new_box src/hittable.rs:132-145: a list of six rects, hit = hit_hittables over them, :43-55) as ONE test: the six
// plane distances are the rects' own t = (k - o_k) / d_k; the line crosses the box between the last near plane (entry) and
// the first far plane (exit).  Closest-of-six = the entry face if its t lies in [t_lo, t_hi], else the exit face — and a ray
// never re-hits the face it starts on (`self_face`, the rects' rule).  Same values as six rect tests except ON an edge,
// where the rect's closed (a, b) interval and the slab comparison may round differently (both are f32).
// Returns t (or NaN) and the face 0..5 in new_box order: z max, z min, y max, y min, x max, x min.
RTW_DEV float box_root(const DPrim* __restrict__ pp, V3 o, V3 d, float t_lo, float t_hi, int self_face, int& face, float* far_root = nullptr) {
    const float4* q = reinterpret_cast<const float4*>(pp);
    const float4 xy = __ldg(q); const float2 z = __ldg(reinterpret_cast<const float2*>(q + 1));
    const float rx = rcp_approx(d.x), ry = rcp_approx(d.y), rz = rcp_approx(d.z);          // (k - o) * rcp(d): rect_root's __fdividef
    const float x0 = (xy.x - o.x) * rx, x1 = (xy.y - o.x) * rx;
    const float y0 = (xy.z - o.y) * ry, y1 = (xy.w - o.y) * ry;
    const float z0 = (z.x - o.z) * rz, z1 = (z.y - o.z) * rz;
    const float nx = fminf(x0, x1), fx = fmaxf(x0, x1), ny = fminf(y0, y1), fy = fmaxf(y0, y1), nz = fminf(z0, z1), fz = fmaxf(z0, z1);
    const float tn = fmaxf(fmaxf(nx, ny), nz), tf = fminf(fminf(fx, fy), fz);
    // entering through the min plane when d > 0; leaving through the max plane when d > 0
    const int fn = tn == nx ? (d.x > 0.f ? 5 : 4) : (tn == ny ? (d.y > 0.f ? 3 : 2) : (d.z > 0.f ? 1 : 0));
    const int ff = tf == fx ? (d.x > 0.f ? 4 : 5) : (tf == fy ? (d.y > 0.f ? 2 : 3) : (d.z > 0.f ? 0 : 1));
    const bool on = tn <= tf;
    const bool n_ok = on && tn >= t_lo && tn <= t_hi && fn != self_face;
    const bool f_ok = on && tf >= t_lo && tf <= t_hi && ff != self_face;
    face = n_ok ? fn : ff;
    if (far_root) *far_root = on ? tf : CUDART_NAN_F;          // the other crossing (ConstantMedium boundary: the second probe's hit)
    return n_ok ? tn : (f_ok ? tf : CUDART_NAN_F);
}

// Any primitive: accepted root in [t_lo, t_hi] or NaN.  `skip` = primitive the ray starts on (-1: none): a planar
// rect cannot be re-hit by a ray leaving it (exact geometry; the reference's f64 gets t ~ 1e-13 < t_min).
// `hit` = the primitive record that describes the hit: `pi` itself, or the face rect of a box.
// object-space ray of the last instance transform used: the faces of a Box share one (ConstantMedium boundary scans)
struct XfCache { int xf; V3 o, d; };
template <int F = FEAT_ALL>
RTW_DEV float prim_root(const DScene& sc, int pi, const TRay& r, float t_lo, float t_hi, int skip, int& hit, float* far_root = nullptr, XfCache* xc = nullptr) {
    const DPrim* pp = sc.prims + pi;
    int4 meta = __ldg(reinterpret_cast<const int4*>(pp) + 3);      // type, mat, xform, 1/dt (moving sphere)
    hit = pi;
    if (!(F & FEAT_RECT) || meta.x <= PRIM_MOVING_SPHERE) return sphere_root(pp, meta.x, __int_as_float(meta.w), r, t_lo, t_hi, pi == skip, far_root);
    if (pi == skip) return CUDART_NAN_F;
    V3 o = r.o, d = r.d;
    if (F & FEAT_RXFORM) {
        if (xc) {                                              // all lanes scan the same boundary: the branch is uniform
            if (xc->xf != meta.z) { xform_ray(sc, meta.z, r, xc->o, xc->d); xc->xf = meta.z; }
            o = xc->o; d = xc->d;
        } else xform_ray(sc, meta.z, r, o, d);                 // xform 0 is the identity: no branch, one instruction stream
    }
    if (meta.x == PRIM_BOX) {
        const int first = __ldg(reinterpret_cast<const int*>(pp) + 6);
        int face;
        const float t = box_root(pp, o, d, t_lo, t_hi, skip >= 0 ? skip - first : -1, face, far_root);
        hit = first + face;
        return t;
    }
    return rect_root(pp, meta.x, o, d, t_lo, t_hi);
}

// AABB::hit (src/aabb.rs:77-103) for the two children of a node, with precomputed 1/d and o/d.  Conservative: closed
// interval, boxes rounded outward, the exit distance padded by 5 ulp (Ize, "Robust BVH ray traversal": 2 ulp for
// correctly rounded 1/d; 3 more for MUFU.RCP) and the rounding of o/d covered per axis (slab_setup).  The f32 test can
// then never cull a primitive the f64 reference would hit; the primitive tests decide.
// Per-ray constants of the slab test.  1/d from MUFU.RCP (relative error <= 2^-23): the SAME inv feeds o*inv, so that error
// is purely relative on every plane distance and is covered by widening Ize's 2-ulp exit padding to 5 ulp.  The ABSOLUTE
// error — the rounding of o/d, 2^-24 |o/d| on a plane distance of ANY size — is covered per axis by folding 4 ulp of
// |o_k/d_k| into the addends of the two FFMAs of that axis: cmn (used with a box's min plane) and cmx (max plane) are
// -o/d moved outward — the near plane's distance shrinks, the far plane's grows, whichever sign d has.  No extra
// instruction per box; and no term shared between axes: a shared slack taken from an axis along which the ray hardly
// moves (|d_k| = 1e-7 |d|: o_k/d_k is 1e7 times the other axes' distances) opened every box on the other two axes, and such
// a ray walked a whole slice of a 1 M-sphere scene (found as multi-millisecond straggler rays in the wavefront trace kernel).
// A direction component of exactly +-0 (axis-parallel test rays; 2^-24 of the scatter draws) would give inv = +-inf and
// plane distances inf - inf = NaN for a box that straddles the origin's coordinate — fmaxf/fminf drop the NaN and the
// test MISSES a box the ray runs inside of (the reference's AABB::hit handles the infinities, src/aabb.rs:77-103).
// Such a component is replaced by +-1e-20: every plane distance stays finite, the origin's side of each slab decides.
RTW_DEV float slab_dir(float d) { return fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d; }
RTW_DEV void slab_setup(V3 o, V3 d, V3& inv, V3& cmn, V3& cmx) {
    inv = mk(rcp_approx(slab_dir(d.x)), rcp_approx(slab_dir(d.y)), rcp_approx(slab_dir(d.z)));
    const V3 oi = mk(o.x * inv.x, o.y * inv.y, o.z * inv.z);
    const float k = 2.384185791015625e-07f;                                  // 4 ulp
    const V3 s = mk(copysignf(k * fabsf(oi.x), inv.x), copysignf(k * fabsf(oi.y), inv.y), copysignf(k * fabsf(oi.z), inv.z));
    cmn = mk(-oi.x - s.x, -oi.y - s.y, -oi.z - s.z);
    cmx = mk(-oi.x + s.x, -oi.y + s.y, -oi.z + s.z);
}
// the same for the wide node test (bvh_wide.h): o/d and the per-axis absolute slack
RTW_DEV void slab_setup_wide(V3 o, V3 d, rtww::WRay& wr) {
    wr.ix = rcp_approx(slab_dir(d.x)); wr.iy = rcp_approx(slab_dir(d.y)); wr.iz = rcp_approx(slab_dir(d.z));
    wr.oix = o.x * wr.ix; wr.oiy = o.y * wr.iy; wr.oiz = o.z * wr.iz;
    const float k = 2.384185791015625e-07f;
    wr.sx = k * fabsf(wr.oix); wr.sy = k * fabsf(wr.oiy); wr.sz = k * fabsf(wr.oiz);
    wr.k = ((wr.ix < 0.0f ? 1u : 0u) | (wr.iy < 0.0f ? 2u : 0u) | (wr.iz < 0.0f ? 4u : 0u)) ^ 7u;
}
RTW_DEV bool slab(float mnx, float mxx, float mny, float mxy, float mnz, float mxz, V3 inv, V3 cmn, V3 cmx, float t_lo, float t_hi, float& t_enter) {
    float x0 = fmaf(mnx, inv.x, cmn.x), x1 = fmaf(mxx, inv.x, cmx.x);
    float y0 = fmaf(mny, inv.y, cmn.y), y1 = fmaf(mxy, inv.y, cmx.y);
    float z0 = fmaf(mnz, inv.z, cmn.z), z1 = fmaf(mxz, inv.z, cmx.z);
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), t_lo));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), t_hi)) * 1.0000006f;
    t_enter = tn;
    return tn <= tf;
}

// The megakernel's variant with ONE slack term shared by the three axes (4 ulp of the largest |o_k/d_k|, folded into the exit
// distance's padding FFMA): two registers fewer than the per-axis form in a kernel that sits at its register cap.  Equally
// conservative; the price is that a ray with |d_k| ~ 1e-7 |d| opens every box on the other two axes — harmless where a warp
// only loses a few hundred node visits to such a ray (scenes below RTW_BIG_MIN under the megakernel), fatal for the
// persistent trace kernel of the wavefront pipeline, where one such ray holds a whole launch (rtw_wavefront.cuh).
// Measured (profiles/r2_z_slab_ab.log): C1 101.65 -> 100.80 ms with the shared term, cornell_box 147.5 -> 148.9, final_scene
// 454.7 -> 456.2: it pays only in the spheres-only variant (F == 0), which is where it is used (-1 = that rule; 0 / 1 force).
#ifndef RTW_SLAB_SHARED
#define RTW_SLAB_SHARED -1
#endif
RTW_DEV void slab_setup_shared(V3 o, V3 d, V3& inv, V3& oi, float& slack) {
    inv = mk(rcp_approx(slab_dir(d.x)), rcp_approx(slab_dir(d.y)), rcp_approx(slab_dir(d.z)));
    oi = mk(o.x * inv.x, o.y * inv.y, o.z * inv.z);
    const float ax = fabsf(d.x) < 1e-20f ? 0.f : fabsf(oi.x), ay = fabsf(d.y) < 1e-20f ? 0.f : fabsf(oi.y), az = fabsf(d.z) < 1e-20f ? 0.f : fabsf(oi.z);
    slack = 2.384185791015625e-07f * fmaxf(fmaxf(ax, ay), az);
}
RTW_DEV bool slab_shared(float mnx, float mxx, float mny, float mxy, float mnz, float mxz, V3 inv, V3 oi, float slack, float t_lo, float t_hi, float& t_enter) {
    float x0 = fmaf(mnx, inv.x, -oi.x), x1 = fmaf(mxx, inv.x, -oi.x);
    float y0 = fmaf(mny, inv.y, -oi.y), y1 = fmaf(mxy, inv.y, -oi.y);
    float z0 = fmaf(mnz, inv.z, -oi.z), z1 = fmaf(mxz, inv.z, -oi.z);
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), t_lo));
    float tf = fmaf(fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), t_hi)), 1.0000006f, slack);
    t_enter = tn;
    return tn <= tf;
}

#define RTW_STACK 64
#ifndef RTW_SPECULATIVE
#define RTW_SPECULATIVE 1
#endif
#define RTW_SENTINEL 0x7fffffff

// Closest surface hit over the BVH (replaces hit_hittables :43-55 + bvh_node_hit :290-306 over the whole world).
// "Speculative while-while" (Aila & Laine 2009): every lane keeps descending inner nodes until ALL lanes of the
// warp hold a leaf (one leaf may be postponed per lane), then the warp intersects leaves together — node visits
// and primitive tests each run with most lanes active instead of interleaving per lane.
#ifdef RTW_INSTRUMENT
__device__ int g_dbg_visits, g_dbg_prims;
#define RTW_DBG_VISIT() (++dbg_visits)
#define RTW_DBG_PRIM() (++dbg_prims)
#else
#define RTW_DBG_VISIT()
#define RTW_DBG_PRIM()
#endif
template <int F = FEAT_ALL>
RTW_DEV void bvh_closest(const DScene& sc, const TRay& r, float t_min, float& t_best, int& prim_best, int skip
#ifdef RTW_INSTRUMENT
                         , int& dbg_visits, int& dbg_prims
#endif
                         ) {
    if (sc.n_bvh_prims == 0) return;
    constexpr bool SH = RTW_SLAB_SHARED < 0 ? (F == 0) : (RTW_SLAB_SHARED != 0);
    V3 inv, cmn, cmx; float slack = 0.f;                 // shared form: cmn = o / d, cmx unused
    if (SH) { slab_setup_shared(r.o, r.d, inv, cmn, slack); cmx = cmn; }
    else slab_setup(r.o, r.d, inv, cmn, cmx);
    int stack[RTW_STACK];
    stack[0] = RTW_SENTINEL;
    int* sp = stack + 1;                          // points at the next free entry
    int node = 0, leaf = 0;                       // leaf >= 0: none postponed
    // Prefetching what is pushed / postponed (prefetch.global.L1 of the far child and of a postponed leaf's record) on scenes
    // beyond the caches — measured and NOT kept: sweep 1 M / 4 M / 16 M 262 / 198 / 160 Mpaths/s with it vs 274 / 207 / 166
    // without, and the dormant code alone costs C1 4 % (105.0 vs 100.9 ms): profiles/r2_g_prefetch.log.  -DRTW_PREFETCH=1 builds it.
#ifndef RTW_PREFETCH
#define RTW_PREFETCH 0
#endif
    const bool big = RTW_PREFETCH && sc.n_nodes > (1 << 16);
    while (node != RTW_SENTINEL) {
        bool searching = true;
        while (node >= 0 && node != RTW_SENTINEL) {
            RTW_DBG_VISIT();
            const float4* np = reinterpret_cast<const float4*>(sc.nodes + node);
            float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            int2 ch = __ldg(reinterpret_cast<const int2*>(np + 3));
            float e0, e1;
            const bool h0 = SH ? slab_shared(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, inv, cmn, slack, t_min, t_best, e0)
                               : slab(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, inv, cmn, cmx, t_min, t_best, e0);
            const bool h1 = SH ? slab_shared(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, inv, cmn, slack, t_min, t_best, e1)
                               : slab(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, inv, cmn, cmx, t_min, t_best, e1);
            // straight-line child selection: nearer hit child next, the other one pushed
            const bool closer1 = e1 < e0;
            const bool second = h1 & (!h0 | closer1);
            const int nearc = second ? ch.y : ch.x, farc = second ? ch.x : ch.y;
            const bool both = h0 && h1, none = !(h0 || h1);
            node = nearc;
            if (both) {
                *sp = farc; ++sp;
                if (big) {
                    const void* pf = farc >= 0 ? (const void*)(sc.nodes + farc) : (const void*)(sc.prims + ((~farc) >> 3));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(pf));
                }
            }
            if (none) { --sp; node = *sp; }
#if RTW_SPECULATIVE
            if (node < 0 && leaf >= 0) {                                                          // postpone first leaf
                searching = false; leaf = node; node = sp[-1]; --sp;
                if (big) asm volatile("prefetch.global.L1 [%0];" :: "l"((const void*)(sc.prims + ((~leaf) >> 3))));
            }
            if (!__any_sync(__activemask(), searching)) break;
#else
            (void)searching;
            if (node < 0) { leaf = node; node = sp[-1]; --sp; break; }
#endif
        }
        while (leaf < 0) {
            int code = ~leaf, first = code >> 3, count = (code & 7) + 1;
#pragma unroll 1                                    // one copy of the primitive tests (the compiler unrolls by 2 when the body is small)
            for (int i = 0; i < count; ++i) {
                RTW_DBG_PRIM();
                int hp;
                float t = prim_root<F>(sc, first + i, r, t_min, t_best, skip, hp);
                if (t == t) { t_best = t; prim_best = hp; }              // not NaN: accepted, t <= t_best
            }
            leaf = node;                                                  // a second leaf was reached meanwhile
            if (node < 0) { node = sp[-1]; --sp; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Closest surface hit over the 8-WIDE COMPRESSED BVH (bvh_wide.h) — scenes that do not fit the caches.
// One node visit = five 16-byte loads and eight quantised slab tests (rtww::wide_node_hits); the children that pass
// form a GROUP (base index + slot mask) from which the ray takes the slot with the highest priority  slot ^ (7 ^ octant)
// (front to back along the ray's octant, no sorting); what is left of the group goes on a stack of 8-byte entries —
// at most one per level of the tree.  Leaves that pass are intersected right away.
// ------------------------------------------------------------------------------------------------
template <int F = FEAT_ALL>
RTW_DEV void bvh8_closest(const DScene& sc, const TRay& r, float t_min, float& t_best, int& prim_best, int skip
#ifdef RTW_INSTRUMENT
                          , int& dbg_visits, int& dbg_prims
#endif
                          ) {
    if (sc.n_bvh_prims <= 0) return;
    rtww::WRay wr;
    slab_setup_wide(r.o, r.d, wr);
    const uint32_t k = wr.k;
    wr.one = 0x3F800000u | ((uint32_t)sc.n_nodes >> 31);           // = 0x3F800000, but opaque to the compiler (see WRay::one)
    uint2 stack[RTW_WIDE_STACK];
    int sp = 0;
    uint32_t base = 0, grp = (1u << 8) | (1u << k);          // the root: a group of one inner slot (slot 0)
    for (;;) {
        if (!(grp & 0xffu)) {
            if (sp == 0) break;
            --sp; base = stack[sp].x; grp = stack[sp].y;
            continue;
        }
        const int j = 31 - __clz(grp & 0xffu);
        grp ^= 1u << j;
        const uint32_t slot = (uint32_t)j ^ k;
        const uint32_t node = base + __popc((grp >> 8) & ((1u << slot) - 1u));
        if (grp & 0xffu) { stack[sp] = make_uint2(base, grp); ++sp; }
        RTW_DBG_VISIT();
        const uint4* np = reinterpret_cast<const uint4*>(sc.wnodes + node);
        const uint4 h = __ldg(np), m = __ldg(np + 1), qa = __ldg(np + 2), qb = __ldg(np + 3), qc = __ldg(np + 4);
        const uint32_t imask = h.w >> 24, lmask = m.z & 0xffu;
        uint32_t hits = rtww::wide_node_hits(*reinterpret_cast<const rtww::W4*>(&h), *reinterpret_cast<const rtww::W4*>(&qa),
                                             *reinterpret_cast<const rtww::W4*>(&qb), *reinterpret_cast<const rtww::W4*>(&qc), wr, t_min, t_best);
        hits &= imask | lmask;
        const uint32_t m16 = rtww::wide_perm16((hits & imask) | ((hits & lmask) << 8), k);
        uint32_t pl = m16 >> 8;
        while (pl) {                                          // leaves, nearest octant first
            const int jj = 31 - __clz(pl);
            pl ^= 1u << jj;
            const uint32_t s = (uint32_t)jj ^ k;
            const int pi = (int)(m.y + __popc(lmask & ((1u << s) - 1u)));
            RTW_DBG_PRIM();
            int hp;
            const float t = prim_root<F>(sc, pi, r, t_min, t_best, skip, hp);
            if (t == t) { t_best = t; prim_best = hp; }
        }
        base = m.x; grp = (imask << 8) | (m16 & 0xffu);
    }
}

// closest surface hit, whichever BVH the scene was committed with (W: compile-time for the render kernel variants)
template <int F = FEAT_ALL, int W = -1>
RTW_DEV void closest_hit(const DScene& sc, const TRay& r, float t_min, float& t_best, int& prim_best, int skip
#ifdef RTW_INSTRUMENT
                         , int& dbg_visits, int& dbg_prims
#endif
                         ) {
    const bool wide = W < 0 ? sc.wnodes != nullptr : W != 0;
#ifdef RTW_INSTRUMENT
    if (wide) bvh8_closest<F>(sc, r, t_min, t_best, prim_best, skip, dbg_visits, dbg_prims);
    else bvh_closest<F>(sc, r, t_min, t_best, prim_best, skip, dbg_visits, dbg_prims);
#else
    if (wide) bvh8_closest<F>(sc, r, t_min, t_best, prim_best, skip);
    else bvh_closest<F>(sc, r, t_min, t_best, prim_best, skip);
#endif
}

// set_face_normal :23-26
RTW_DEV void set_face_normal(V3 dir, V3 outward, V3& normal, int& front) {
    front = dot(dir, outward) < 0.0f;
    normal = front ? outward : -outward;
}

// Fill the HitRecord for the accepted (prim, t): sphere_hit :275-287, rect_hit :322-329, then the wrapper chain
// (Translate :236-239, hit_rotate_y :398-410) with the reference's nested set_face_normal calls replayed literally.
template <int F = FEAT_ALL>
RTW_DEV void finalize_hit(const DScene& sc, int pi, float t, const TRay& r, bool want_uv, HitRec& rec) {
    const DPrim* pp = sc.prims + pi;
    int4 meta = __ldg(reinterpret_cast<const int4*>(pp) + 3);
    rec.t = t; rec.mat = meta.y; rec.u = 0.0f; rec.v = 0.0f;
    Ray wr; wr.o = r.o; wr.d = r.d;
    rec.p = ray_at(wr, t);
    V3 outward_obj, d_obj = r.d;
    const int xf = (F & FEAT_XFORM) ? meta.z : 0;
    float mc = 1.0f, ms = 0.0f;
    if (xf) { float4 m = __ldg(reinterpret_cast<const float4*>(sc.xforms + xf)); mc = m.x; ms = m.y; d_obj = mk(mc * r.d.x - ms * r.d.z, r.d.y, ms * r.d.x + mc * r.d.z); }
    if (!(F & FEAT_RECT) || meta.x <= PRIM_MOVING_SPHERE) {
        double cx, cy, cz, rad;
        load_prim_center(pp, meta.x, __int_as_float(meta.w), r.time, cx, cy, cz, rad);
        // one Newton step of f(t) = a t^2 + 2 half_b t + c in f64: t, the hit point and the normal then carry
        // the reference's precision (an f32 t alone leaves |t d| * 1e-7 / r ~ 1e-5 of error on small far spheres)
        double ocx = r.gox() - cx, ocy = r.goy() - cy, ocz = r.goz() - cz;
        double half_b = ocx * r.gdx() + ocy * r.gdy() + ocz * r.gdz();
        double c = ocx * ocx + ocy * ocy + ocz * ocz - rad * rad;
        double td = (double)t;
        double f = (r.ga() * td + 2.0 * half_b) * td + c, fp = 2.0 * (r.ga() * td + half_b);
        td -= (double)__fdividef((float)f, (float)fp);
        rec.t = (float)td;
        double px = fma(td, r.gdx(), r.gox()), py = fma(td, r.gdy(), r.goy()), pz = fma(td, r.gdz(), r.goz());
        rec.p = mk((float)px, (float)py, (float)pz);
        float inv_r = rcp_approx((float)rad);
        V3 ow = mk((float)(px - cx) * inv_r, (float)(py - cy) * inv_r, (float)(pz - cz) * inv_r);
        outward_obj = xf ? mk(mc * ow.x - ms * ow.z, ow.y, ms * ow.x + mc * ow.z) : ow;
        if (want_uv) sphere_uv(outward_obj, rec.u, rec.v);
    } else {
        V3 o = r.o, d = r.d;
        if (xf) xform_ray(sc, xf, r, o, d);
        const float4* q = reinterpret_cast<const float4*>(pp);
        float4 ab = __ldg(q);
        float a, b;
        if (meta.x == PRIM_XY) { a = o.x + t * d.x; b = o.y + t * d.y; outward_obj = mk(0.f, 0.f, 1.f); }
        else if (meta.x == PRIM_XZ) { a = o.x + t * d.x; b = o.z + t * d.z; outward_obj = mk(0.f, 1.f, 0.f); }
        else { a = o.y + t * d.y; b = o.z + t * d.z; outward_obj = mk(1.f, 0.f, 0.f); }
        rec.u = __fdividef(a - ab.x, ab.y - ab.x);
        rec.v = __fdividef(b - ab.z, ab.w - ab.z);
    }
    set_face_normal(d_obj, outward_obj, rec.normal, rec.front);
    if (xf) {
        const DXform* x = sc.xforms + xf;
        int n_ops = __ldg(&x->n_ops);
        V3 nrm = rec.normal; int front = rec.front;
        for (int j = n_ops - 1; j >= 0; --j) {
            float4 op = __ldg(reinterpret_cast<const float4*>(&x->ops[j]));
            // normal into the space outside op j (inverse rotation; identity for Translate)
            nrm = mk(op.x * nrm.x + op.y * nrm.z, nrm.y, -op.y * nrm.x + op.x * nrm.z);
            // direction on the INSIDE of op j (object-space ray of hit_rotate_y :409; moved_ray of Translate :238)
            V3 dj = mk(op.z * r.d.x - op.w * r.d.z, r.d.y, op.w * r.d.x + op.z * r.d.z);
            V3 outn = nrm;
            set_face_normal(dj, outn, nrm, front);
        }
        rec.normal = nrm; rec.front = front;
    }
}

// hit_constant_medium :417-473 for medium m, with the current closest surface hit as t_max.
template <class R>
RTW_DEV bool medium_hit(const DScene& sc, int mi, const TRay& r, float t_min, float t_max, R& g, float& t_out, int& mat_out) {
    int4 md = __ldg(reinterpret_cast<const int4*>(sc.media + mi));
    const float inf = CUDART_INF_F;
    // boundary.hit(ray, -inf, inf) :422, then boundary.hit(ray, rec1.t + 0.0001, inf) :423 — two closest-so-far scans
    // over the boundary prims (hit_hittables :43-55) sharing ONE inlined copy of the primitive test
    // A boundary made of ONE sphere (final_scene's two media): its second crossing comes out of the same discriminant
    // — the first probe returns the near root, the second probe can only return the far one — so one test serves both.
    const bool one_sphere = md.y == 1 && __ldg(reinterpret_cast<const int*>(sc.prims + md.x) + 12) <= PRIM_MOVING_SPHERE;
    float t1 = 0.f, t2 = 0.f, lo = -inf;
    if (one_sphere) {
        // straight-line path (final_scene: both media, every segment of every path): one sphere test, no scan loops, no root cache
        const DPrim* pp = sc.prims + md.x;
        const int4 meta = __ldg(reinterpret_cast<const int4*>(pp) + 3);
        float far_root = CUDART_NAN_F;
        t1 = sphere_root(pp, meta.x, __int_as_float(meta.w), r, -inf, inf, false, &far_root);       // the near crossing, over (-inf, inf)
        if (!(t1 == t1)) return false;
        // in f32 the +0.0001 vanishes once |t1| > 2048 (r = 5000 fog sphere): keep the probe strictly beyond t1
        lo = t1 + 0.0001f;
        if (!(lo > t1)) lo = nextafterf(t1, inf);
        if (!(far_root >= lo)) return false;
        t2 = far_root;
    } else {
    XfCache xc; xc.xf = -1; xc.o = r.o; xc.d = r.d;        // 6 faces x 2 probes of a rotated Box: one transform instead of 12
    // Small boundaries (a Box: 6 faces) are intersected ONCE: the first scan keeps every face's root, the second scan
    // (same faces, range [t1 + 0.0001, inf)) re-reads them — the same values hit_hittables would compute again.
    // A sphere among those faces returns its NEAR root to the (-inf, inf) probe; the second probe may need the FAR one
    // (boundary.hit(rec1.t + 0.0001, inf) on a BvhNode / list of spheres), so both roots are kept.
    float ts[8], tfar[8];
    const bool keep = md.y <= 8;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        float hi = inf, found = CUDART_NAN_F;
#pragma unroll 1
        for (int i = 0; i < md.y; ++i) {
            float t;
            if (keep && pass == 1) { t = ts[i]; if (!(t >= lo)) t = tfar[i]; }
            else {
                float fr = CUDART_NAN_F;                             // stays NaN for rects
                int hp;                                              // (a Box boundary is ONE record: its slab test yields both crossings)
                t = prim_root(sc, md.x + i, r, keep ? -inf : lo, keep ? inf : hi, -1, hp, &fr, &xc);
                if (keep) { ts[i] = t; tfar[i] = fr; }
            }
            if (t >= lo && t <= hi) { hi = t; found = t; }          // closest-so-far (hit_hittables :43-55); NaN fails
        }
        if (!(found == found)) return false;
        if (pass == 0) {
            t1 = found;
            lo = t1 + 0.0001f;
            if (!(lo > t1)) lo = nextafterf(t1, inf);
        } else t2 = found;
    }
    }
    if (t1 < t_min) t1 = t_min;
    if (t2 > t_max) t2 = t_max;
    if (t1 >= t2) return false;
    if (t1 < 0.0f) t1 = 0.0f;
    float ray_length = sqrt_approx(length_squared(r.d));
    float distance_inside_boundary = (t2 - t1) * ray_length;
    float hit_distance = __int_as_float(md.z) * logf(g.next());                            // :446 — the draw
    if (hit_distance > distance_inside_boundary) return false;
    t_out = t1 + __fdividef(hit_distance, ray_length);
    mat_out = md.w;
    return true;
}

// World closest hit = surfaces through the BVH, then the media in list order (media_deferred order of the oracle).
template <class R>
RTW_DEV bool world_hit(const DScene& sc, const TRay& r, float t_min, float t_max, R& g, bool want_uv, HitRec& rec) {
    float t_best = t_max; int prim_best = -1;
#ifdef RTW_INSTRUMENT
    int dv = 0, dp = 0;
    closest_hit(sc, r, t_min, t_best, prim_best, -1, dv, dp);
#else
    closest_hit(sc, r, t_min, t_best, prim_best, -1);
#endif
    int med_mat = -1; float med_t = 0.0f;
    for (int m = 0; m < sc.n_media; ++m) {
        float t; int mat;
        if (medium_hit(sc, m, r, t_min, t_best, g, t, mat)) { t_best = t; med_t = t; med_mat = mat; prim_best = -2; }
    }
    if (prim_best == -1) return false;
    if (prim_best == -2) {                                                                 // :460-464
        Ray wr; wr.o = r.o; wr.d = r.d;
        rec.t = med_t; rec.p = ray_at(wr, med_t); rec.normal = mk(1.f, 0.f, 0.f); rec.front = 1; rec.mat = med_mat;
        rec.u = 0.f; rec.v = 0.f;
        return true;
    }
    finalize_hit(sc, prim_best, t_best, r, want_uv, rec);
    return true;
}

// ------------------------------------------------------------------------------------------------
// src/material.rs — emitted (:25-34), scatter (:15-23, :36-94)
// ------------------------------------------------------------------------------------------------
RTW_DEV float reflectance(float cosine, float ref_idx) {                                    // :89-94
    float r0 = __fdividef(1.0f - ref_idx, 1.0f + ref_idx);
    r0 = r0 * r0;
    float x = 1.0f - cosine, x2 = x * x;
    return r0 + (1.0f - r0) * (x2 * x2 * x);                                               // powf(5.0)
}

struct DMatRec { float r, g, b, param; int kind, tex; };
RTW_DEV DMatRec load_mat(const DScene& sc, int mat) {
    const float4* mp = reinterpret_cast<const float4*>(sc.mats + mat);
    float4 a = __ldg(mp); int4 b = __ldg(reinterpret_cast<const int4*>(mp + 1));
    DMatRec m; m.r = a.x; m.g = a.y; m.b = a.z; m.param = a.w; m.kind = b.x; m.tex = b.y;
    return m;
}
template <int F = FEAT_ALL>
RTW_DEV bool mat_needs_uv(const DScene& sc, const DMatRec& m) {
    if (!(F & FEAT_IMAGE) || m.tex < 0) return false;
    return __float_as_int(__ldg(&reinterpret_cast<const float4*>(sc.texs + m.tex)[1].w)) == TEX_IMAGE;
}
template <int F = FEAT_ALL>
RTW_DEV V3 mat_color(const DScene& sc, const DMatRec& m, const HitRec& rec) {
    if (m.tex < 0) return mk(m.r, m.g, m.b);
    return texture_value<F>(sc, m.tex, rec.u, rec.v, rec.p);
}

// Returns true when a scattered ray exists.  `emitted` is always written.
// Lambertian, Metal and Isotropic all start with the same unit-ball rejection loop (src/math.rs:51-58): it is
// hoisted to ONE call site so the lanes of a warp run it — and the Philox blocks behind it — together.
RTW_DEV bool mat_needs_ball(int kind) { return kind != MAT_DIFFUSE_LIGHT && kind != MAT_DIELECTRIC; }
template <class R, int F = FEAT_ALL>
RTW_DEV bool scatter(const DScene& sc, const DMatRec& m, const Ray& ray, const HitRec& rec, R& g, Ray& scattered, V3& attenuation, V3& emitted,
                     const V3* ball_in = nullptr) {
    emitted = mk(0.f, 0.f, 0.f);
    scattered.o = rec.p; scattered.time = ray.time;
    const int kind = m.kind;
    if (kind == MAT_DIFFUSE_LIGHT) {                                                       // :20, :25-34 (both faces emit)
        emitted = mat_color<F>(sc, m, rec);
        attenuation = mk(0.f, 0.f, 0.f);
        scattered.d = mk(0.f, 0.f, 0.f);
        return false;
    }
    if (kind == MAT_DIELECTRIC) {                                                          // :62-82
        // f64 island: near the critical angle 1 - |r_perp|^2 (src/math.rs:114) cancels to ~1e-5 and an f32 cos(theta)
        // (1e-7 absolute) would leave 1e-4 on the refracted direction.  B200: FP64 at half rate, 8 % of the hits.
        attenuation = mk(1.f, 1.f, 1.f);
        // 1 / ir: MUFU.RCP + one f64 Newton step (error 2^-46) instead of the f64 division sequence
        const double ir = (double)m.param, ir_r = (double)rcp_approx(m.param);
        const double ratio = rec.front ? ir_r * (2.0 - ir * ir_r) : ir;
        const double dx = ray.d.x, dy = ray.d.y, dz = ray.d.z, nx = rec.normal.x, ny = rec.normal.y, nz = rec.normal.z;
        const double dd = dx * dx + dy * dy + dz * dz;
        double inv = (double)rsqrtf((float)dd);
        inv = inv * (1.5 - 0.5 * dd * inv * inv);                                          // one Newton step: 2e-7 -> 1e-13
        const double ux = dx * inv, uy = dy * inv, uz = dz * inv;                          // unit_direction :66
        const double cos_theta = fmin(-(ux * nx + uy * ny + uz * nz), 1.0);                // :67
        const double sin2 = 1.0 - cos_theta * cos_theta;
        const bool cannot_refract = ratio * ratio * sin2 > 1.0;                            // ratio * sin_theta > 1 :70
        if (cannot_refract || reflectance((float)cos_theta, (float)ratio) > g.next()) {   // short-circuit draw :72
            const double k2 = 2.0 * (ux * nx + uy * ny + uz * nz);                         // reflect :106-108
            scattered.d = mk((float)(ux - k2 * nx), (float)(uy - k2 * ny), (float)(uz - k2 * nz));
        } else {                                                                           // refract :110-117
            const double px = ratio * (ux + cos_theta * nx), py = ratio * (uy + cos_theta * ny), pz = ratio * (uz + cos_theta * nz);
            const double k = 1.0 - (px * px + py * py + pz * pz);
            const double par = -(double)sqrt_approx((float)fabs(k));
            scattered.d = mk((float)(px + par * nx), (float)(py + par * ny), (float)(pz + par * nz));
        }
        return true;
    }
    const V3 ball = ball_in ? *ball_in : random_in_unit_sphere(g);                         // :37 / :52 / :85 (drawn by the warp: coop_unit_sphere)
    if (kind == MAT_LAMBERTIAN) {                                                          // :36-48
        V3 dir = rec.normal + normalize(ball);
        if (near_zero(dir)) dir = rec.normal;
        scattered.d = dir;
        attenuation = mat_color<F>(sc, m, rec);
        return true;
    }
    if (kind == MAT_METAL) {                                                               // :50-60
        V3 reflected = reflect(normalize(ray.d), rec.normal);
        scattered.d = reflected + m.param * ball;
        attenuation = mk(m.r, m.g, m.b);
        return dot(scattered.d, rec.normal) > 0.0f;
    }
    scattered.d = ball;                                                                    // Isotropic :84-87
    attenuation = mat_color<F>(sc, m, rec);
    return true;
}

// ------------------------------------------------------------------------------------------------
// src/main.rs:19-38 — ray_color, unrolled into an iterative throughput/radiance pair:
//   L += T * emitted;  T *= attenuation;  miss: L += T * background;  depth exhausted: nothing added (:21-23)
// ------------------------------------------------------------------------------------------------
struct PathState {
    Ray ray;
    V3 T;                   // throughput; the radiance terms T*emitted / T*background are handed to the caller
    int segment;            // segments traced so far (bounce id of the next one = segment + 1)
    int last_prim;          // primitive the current ray starts on (-1: camera / medium scatter)
#ifdef RTW_INSTRUMENT
    int dbg_visits, dbg_prims;
#endif
    PhiloxRng rng;
};

// Work unit -> (tile, sample range).  Units are handed out in index order from one counter: first the big phase-A
// units, then the small phase-B ones (guided self-scheduling: the frame ends on short units, so the last warp finishes
// at most one SHORT unit after the others instead of one long one).
RTW_DEV void decode_unit(const DParams& prm, unsigned unit, int& tile, int& s0, int& s1) {
    if (unit < prm.n_units_a) {
        tile = (int)(unit / (unsigned)prm.chunks);
        s0 = (int)(unit % (unsigned)prm.chunks) * prm.chunk_spp;
        s1 = min(prm.spp_a, s0 + prm.chunk_spp);
    } else {
        const unsigned u = unit - prm.n_units_a;
        tile = (int)(u / (unsigned)prm.chunks_b);
        s0 = prm.spp_a + (int)(u % (unsigned)prm.chunks_b) * prm.chunk_spp_b;
        s1 = min(prm.spp, s0 + prm.chunk_spp_b);
    }
    s0 += prm.first_sample; s1 += prm.first_sample;      // progressive pass: the Philox sample coordinate is absolute
}

// Finished tile (32 pixels x rgb in shared memory, 16-byte aligned) -> framebuffer; row 0 of the image = top
// (y = H-1 of src/main.rs:591).  When several units or GPUs contribute to a pixel the tile is ADDED: a full-width tile
// row is 24 contiguous floats, 16-byte aligned whenever the image width is a multiple of 4, so 24 lanes issue one
// red.v4.f32 each (6 per row) instead of 96 scalar atomics — over NVLink that is 4x fewer packets into GPU 0.
RTW_DEV void flush_tile(const DParams& prm, float* __restrict__ fb, const float* acc, int tx, int ty, int tw, int lane) {
    if (prm.accumulate && tw == 8 && (prm.width & 3) == 0 && (reinterpret_cast<uintptr_t>(fb) & 15) == 0) {
        const int row = lane / 6, q = lane - row * 6, y = ty * 4 + row;
        if (lane < 24 && y < prm.height) {
            float* dst = fb + ((size_t)(prm.height - 1 - y) * prm.width + tx * 8) * 3 + q * 4;
            const float4 v = *reinterpret_cast<const float4*>(acc + row * 24 + q * 4);
            asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
        return;
    }
    const int x = tx * 8 + (lane & 7), y = ty * 4 + (lane >> 3);
    if (x < prm.width && y < prm.height) {
        float* dst = fb + ((size_t)(prm.height - 1 - y) * prm.width + x) * 3;
        const float r = acc[lane * 3], g = acc[lane * 3 + 1], b = acc[lane * 3 + 2];
        if (prm.accumulate) { atomicAdd_system(dst, r); atomicAdd_system(dst + 1, g); atomicAdd_system(dst + 2, b); }
        else { dst[0] = r; dst[1] = g; dst[2] = b; }
    }
}

RTW_DEV void path_begin(const DCamera& cam, const DParams& prm, int x, int y, int s, PathState& ps, bool bound = false) {   // :517-520
    if (!bound) ps.rng.bind(prm);
    ps.rng.start((uint32_t)(y * prm.width + x), (uint32_t)s);
    ps.ray = camera_ray_loop<true>(cam, 0.f, 0.f, (float)x, (float)y, (float)prm.width - 1.0f, (float)prm.height - 1.0f, ps.rng);
    ps.T = mk(1.f, 1.f, 1.f);
    ps.segment = 0; ps.last_prim = -1;
}

// Second half of one level of ray_color: media (list order, after the surfaces), miss -> background, otherwise
// hit record + Material::emitted/scatter.  `tr`/(t_best, prim_best) = the segment's ray and its closest surface hit.
// Returns true while the path continues; `add` = this level's radiance term (often 0).
// COOP: called by ALL 32 lanes of the warp (`work` = this lane has a segment to finish); the unit-ball sample of the
// lanes that need one is then drawn by the whole warp (coop_unit_sphere) — BEFORE the hit record is built, while few
// values are live; the f64 copies of the ray are re-derived afterwards instead of being kept across the sampling loop.
// !COOP: independent threads (parity hooks).
RTW_DEV float opaque(float v) { asm volatile("" : "+f"(v)); return v; }
template <int F = FEAT_ALL, bool COOP = false>
RTW_DEV bool path_finish(const DScene& sc, const DParams& prm, PathState& ps, const TRay& tr, float t_best, int prim_best, V3& add,
                         bool work = true, int lane = 0, int* scr = nullptr) {
    add = mk(0.f, 0.f, 0.f);
    int med_mat = -1, mat = -1;
    if (work) {
        if (F & FEAT_MEDIA) for (int mi = 0; mi < sc.n_media; ++mi) {
            float t; int mm;
            if (medium_hit(sc, mi, tr, prm.t_min, t_best, ps.rng, t, mm)) { t_best = t; med_mat = mm; prim_best = -2; }
        }
        if (prim_best == -1) add = ps.T * mk(prm.bg_r, prm.bg_g, prm.bg_b);                // :37
        else mat = ((F & FEAT_MEDIA) && prim_best == -2) ? med_mat : __ldg(&sc.prims[prim_best].mat);
    }
    V3 ball = mk(0.f, 0.f, 0.f);
    if (COOP) ball = coop_unit_sphere(mat >= 0 && mat_needs_ball(__ldg(&sc.mats[mat].kind)), ps.rng, lane, scr);
    if (mat < 0) return false;
    HitRec rec;
    const DMatRec m = load_mat(sc, mat);                                                   // :26
    if ((F & FEAT_MEDIA) && prim_best == -2) {
        rec.t = t_best; rec.p = ray_at(ps.ray, t_best); rec.normal = mk(1.f, 0.f, 0.f); rec.front = 1; rec.mat = med_mat; rec.u = 0.f; rec.v = 0.f;
    } else if (COOP) {
        Ray rr; rr.o = mk(opaque(ps.ray.o.x), opaque(ps.ray.o.y), opaque(ps.ray.o.z)); rr.d = mk(opaque(ps.ray.d.x), opaque(ps.ray.d.y), opaque(ps.ray.d.z));
        rr.time = ps.ray.time;
        finalize_hit<F>(sc, prim_best, t_best, make_tray(rr), mat_needs_uv<F>(sc, m), rec);
    } else finalize_hit<F>(sc, prim_best, t_best, tr, mat_needs_uv<F>(sc, m), rec);
    Ray scattered; V3 att, em;
    bool cont = scatter<PhiloxRng, F>(sc, m, ps.ray, rec, ps.rng, scattered, att, em, COOP ? &ball : nullptr);   // :28-33
    add = ps.T * em;
    if (!cont) return false;
    ps.T = ps.T * att;
    ps.ray = scattered;
    ps.last_prim = prim_best >= 0 ? prim_best : -1;
    return true;
}

// One level of ray_color: closest hit through the BVH, then path_finish.
RTW_DEV bool path_step(const DScene& sc, const DParams& prm, PathState& ps, V3& add) {
    add = mk(0.f, 0.f, 0.f);
    if (ps.segment >= prm.max_depth) return false;                                         // :21-23
    ps.segment++;
    ps.rng.set_bounce((uint32_t)ps.segment);
    TRay tr = make_tray(ps.ray);
    float t_best = CUDART_INF_F; int prim_best = -1;
#ifdef RTW_INSTRUMENT
    closest_hit(sc, tr, prm.t_min, t_best, prim_best, ps.last_prim, ps.dbg_visits, ps.dbg_prims);
#else
    closest_hit(sc, tr, prm.t_min, t_best, prim_best, ps.last_prim);                       // :25
#endif
    return path_finish(sc, prm, ps, tr, t_best, prim_best, add);
}

// ------------------------------------------------------------------------------------------------
// Tile culling for PRIMARY rays.  All camera rays of one work unit leave a lens disk of radius R and cross an 8x4-pixel
// window of the focus plane: origin and direction are confined to two small boxes (interval arithmetic over pixel
// range, jitter and lens offset; src/camera.rs:58-66, src/main.rs:517-518).  One conservative interval-slab walk of
// the BVH per unit collects every primitive such a ray could touch (moving spheres: node boxes span the shutter);
// the unit's thousands of primary rays then test that short list instead of traversing — dense, same trip count in
// every lane.  Exactness: the list is a superset of what any of those rays can hit, the primitive tests are the same.
// ------------------------------------------------------------------------------------------------
#define RTW_TILE_LIST 48

// Lens box L (where every ray starts, tau = 0) and focus-plane window Wd (where every ray is at tau = 1; it does not
// depend on the lens offset: o + d = origin + llc_rel + s h + t v).  A ray of the tile is at (1 - tau) l + tau w for some
// l in L, w in Wd — keeping that correlation makes the bundle as thin as the pixel window near the focus plane instead
// of one lens diameter wide.
struct RayBounds { float ll[3], lh[3], wl[3], wh[3]; };

RTW_DEV RayBounds tile_ray_bounds(const DCamera& c, const DParams& prm, int x0, int y0, int tw, int th) {
    // (2-ulp divisions: the window is padded by 2e-5 relative below)
    const float iw = rcp_approx((float)prm.width - 1.0f), ih = rcp_approx((float)prm.height - 1.0f);
    const float s_lo = (float)x0 * iw, s_hi = (float)(x0 + tw) * iw;
    const float t_lo = (float)y0 * ih, t_hi = (float)(y0 + th) * ih;
    const float sc_ = 0.5f * (s_lo + s_hi), ds = 0.5f * (s_hi - s_lo), tc = 0.5f * (t_lo + t_hi), dt = 0.5f * (t_hi - t_lo);
    const float R = fabsf(c.lens_radius);
    const float o[3] = {c.ox, c.oy, c.oz}, l[3] = {c.lx, c.ly, c.lz}, h[3] = {c.hx, c.hy, c.hz}, v[3] = {c.vx, c.vy, c.vz};
    const float u[3] = {c.ux, c.uy, c.uz}, w[3] = {c.wx, c.wy, c.wz};
    RayBounds b;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float off = R * (fabsf(u[a]) + fabsf(w[a]));
        const float wc = o[a] + l[a] + sc_ * h[a] + tc * v[a];
        const float wh_ = fabsf(h[a]) * ds + fabsf(v[a]) * dt;
        const float pad = 2e-5f * (fabsf(o[a]) + fabsf(wc) + off + wh_) + 1e-6f;      // f32 rounding of the real rays
        b.ll[a] = o[a] - off - pad; b.lh[a] = o[a] + off + pad;
        b.wl[a] = wc - wh_ - pad;   b.wh[a] = wc + wh_ + pad;
    }
    return b;
}

// tau-interval on which  x + tau * dx <= hi  (le = true)  or  x + tau * dx >= lo  (le = false), intersected into [a, b]
RTW_DEV void clip_linear(float x, float dx, float bound, bool le, float& a, float& b) {
    const float n = bound - x;
    if (dx == 0.0f) { if (le ? (n < 0.0f) : (n > 0.0f)) { a = 1.0f; b = 0.0f; } return; }
    const float t = __fdividef(n, dx);              // 2 ulp, inside the 1e-5 slack of bounds_hit_box
    if ((dx > 0.0f) == le) b = fminf(b, t); else a = fmaxf(a, t);
}

// Could ANY ray of the tile hit the box for some tau >= t_min?  Conservative (per-axis intervals).
RTW_DEV bool bounds_hit_box(const RayBounds& rb, float mnx, float mxx, float mny, float mxy, float mnz, float mxz, float t_min) {
    const float mn[3] = {mnx, mny, mnz}, mx[3] = {mxx, mxy, mxz};
    // region A: tau in [t_min, 1]: position in [ll + tau (wl - ll), lh + tau (wh - lh)]
    float a0 = t_min, a1 = 1.0f;
    // region B: tau in [1, inf):  (1 - tau) <= 0 swaps the lens bounds: [lh + tau (wl - lh), ll + tau (wh - ll)]
    float b0 = 1.0f, b1 = CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        clip_linear(rb.ll[k], rb.wl[k] - rb.ll[k], mx[k], true, a0, a1);
        clip_linear(rb.lh[k], rb.wh[k] - rb.lh[k], mn[k], false, a0, a1);
        clip_linear(rb.lh[k], rb.wl[k] - rb.lh[k], mx[k], true, b0, b1);
        clip_linear(rb.ll[k], rb.wh[k] - rb.ll[k], mn[k], false, b0, b1);
    }
    const float sa = 1e-5f * (fabsf(a0) + fabsf(a1)) + 1e-6f, sb = 1e-5f * (fabsf(b0) + (b1 < CUDART_INF_F ? fabsf(b1) : 0.0f)) + 1e-6f;
    return (a0 <= a1 + sa) || (b0 <= b1 + sb);
}

// Warp-cooperative, breadth-first walk of the BVH with the tile's ray bundle: each lane tests one (node, child) pair
// per round, surviving inner children form the next frontier, surviving leaves append their primitives to `list`.
// `scratch` = 2 x 64 frontier ints + 2 counters of per-warp shared memory (the ring, empty at unit start).
// Returns the number of candidates, or -1 when the list (RTW_TILE_LIST) or a frontier (64) overflows — the caller
// then traverses the BVH per ray as usual.
RTW_DEV int build_tile_list(const DScene& sc, const RayBounds& rb, float t_min, int* list, int* scratch, int lane) {
    if (sc.n_bvh_prims == 0) return 0;
    int* front[2] = {scratch, scratch + 64};
    int* cnt = scratch + 128;                    // [0] next frontier size, [1] list size
    int cur = 0, n_front = 1;
    if (lane == 0) { front[0][0] = 0; cnt[0] = 0; cnt[1] = 0; }
    __syncwarp();
    bool overflow = false;
    while (n_front > 0) {
        const int n_tests = 2 * n_front;
        for (int base = 0; base < n_tests; base += 32) {
            const int i = base + lane;
            if (i < n_tests) {
                const int node = front[cur][i >> 1], c = i & 1;
                const float4* np = reinterpret_cast<const float4*>(sc.nodes + node);
                const float4 bx = __ldg(np + c), n2 = __ldg(np + 2);
                const int2 ch = __ldg(reinterpret_cast<const int2*>(np + 3));
                const int child = c ? ch.y : ch.x;
                const bool dup = c && ch.y == ch.x;                      // single-leaf root fills both slots
                if (!dup && bounds_hit_box(rb, bx.x, bx.y, bx.z, bx.w, c ? n2.z : n2.x, c ? n2.w : n2.y, t_min)) {
                    if (child >= 0) {
                        const int pos = atomicAdd(&cnt[0], 1);
                        if (pos < 64) front[cur ^ 1][pos] = child;
                    } else {
                        const int code = ~child, first = code >> 3, count = (code & 7) + 1;
                        const int pos = atomicAdd(&cnt[1], count);
                        if (pos + count <= RTW_TILE_LIST) for (int k = 0; k < count; ++k) list[pos + k] = first + k;
                    }
                }
            }
        }
        __syncwarp();
        n_front = cnt[0];
        overflow |= n_front > 64 || cnt[1] > RTW_TILE_LIST;
        __syncwarp();
        if (lane == 0) cnt[0] = 0;
        cur ^= 1;
        if (overflow) break;
        __syncwarp();
    }
    __syncwarp();
    const int n = cnt[1];
    __syncwarp();
    return overflow ? -1 : n;
}

// The same walk over the 8-wide compressed BVH: one (node, slot) pair per lane and round; child boxes are decoded from
// the quantised planes with directed rounding (the decoded box contains the stored one).
RTW_DEV int build_tile_list_wide(const DScene& sc, const RayBounds& rb, float t_min, int* list, int* scratch, int lane) {
    if (sc.n_bvh_prims == 0) return 0;
    int* front[2] = {scratch, scratch + 64};
    int* cnt = scratch + 128;
    int cur = 0, n_front = 1;
    if (lane == 0) { front[0][0] = 0; cnt[0] = 0; cnt[1] = 0; }
    __syncwarp();
    bool overflow = false;
    while (n_front > 0) {
        const int n_tests = 8 * n_front;
        for (int b0 = 0; b0 < n_tests; b0 += 32) {
            const int i = b0 + lane;
            if (i < n_tests) {
                const int node = front[cur][i >> 3], s = i & 7;
                const DWNode* wn = sc.wnodes + node;
                const uint4 h = __ldg(reinterpret_cast<const uint4*>(wn)), m = __ldg(reinterpret_cast<const uint4*>(wn) + 1);
                const uint32_t imask = h.w >> 24, lmask = m.z & 0xffu;
                if (((imask | lmask) >> s) & 1u) {
                    const uint8_t* q = reinterpret_cast<const uint8_t*>(wn) + 32;
                    const float sx = __uint_as_float((h.w & 0xffu) << 23), sy = __uint_as_float(((h.w >> 8) & 0xffu) << 23), sz = __uint_as_float(((h.w >> 16) & 0xffu) << 23);
                    const float px = __uint_as_float(h.x), py = __uint_as_float(h.y), pz = __uint_as_float(h.z);
                    const float mnx = __fadd_rd(px, (float)__ldg(q + s) * sx), mny = __fadd_rd(py, (float)__ldg(q + 8 + s) * sy), mnz = __fadd_rd(pz, (float)__ldg(q + 16 + s) * sz);
                    const float mxx = __fadd_ru(px, (float)__ldg(q + 24 + s) * sx), mxy = __fadd_ru(py, (float)__ldg(q + 32 + s) * sy), mxz = __fadd_ru(pz, (float)__ldg(q + 40 + s) * sz);
                    if (bounds_hit_box(rb, mnx, mxx, mny, mxy, mnz, mxz, t_min)) {
                        const uint32_t below = (1u << s) - 1u;
                        if ((imask >> s) & 1u) {
                            const int pos = atomicAdd(&cnt[0], 1);
                            if (pos < 64) front[cur ^ 1][pos] = (int)(m.x + __popc(imask & below));
                        } else {
                            const int pos = atomicAdd(&cnt[1], 1);
                            if (pos < RTW_TILE_LIST) list[pos] = (int)(m.y + __popc(lmask & below));
                        }
                    }
                }
            }
        }
        __syncwarp();
        n_front = cnt[0];
        overflow |= n_front > 64 || cnt[1] > RTW_TILE_LIST;
        __syncwarp();
        if (lane == 0) cnt[0] = 0;
        cur ^= 1;
        if (overflow) break;
        __syncwarp();
    }
    __syncwarp();
    const int n = cnt[1];
    __syncwarp();
    return overflow ? -1 : n;
}

}  // namespace rtwd

#endif
