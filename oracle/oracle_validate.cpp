// TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's motion-validation path.
//
// Plain scalar C++ (no SIMD, no CUDA): the checker the GPU engine is compared against when the compiled
// reference (oracle/_ref) is not the one in use, and the source of the instrumented figures
// (W_ref flops, per-edge minimum clearance) that the reference itself does not report.
// Pinned against oracle/_ref (the unmodified reference compiled here) by tests/test_oracle.py and
// against the golden vectors in tests/golden/ that were generated from it.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
//
// Reference followed (paths under /root/reference/external/vamp/src/impl/vamp/):
//   planning/validate.hh:23-77      rake interpolation, n = max(ceil(dist/8*res),1), back-stepping
//   vector/avx.hh:455-549           cephes sin with ROUNDED octant index
//   vector/interface.hh:451-459     cos(x) = sin(wrap(x + pi/2))
//   vector/avx.hh:442-452           hsum order used by l2_norm
//   collision/validity.hh:46-150    ordered scan per shape class with `break` on the bound
//   collision/sphere_*.hh           the primitive tests
//   collision/shapes.hh:52-67,165-189,236-239   min_distance of each shape
//   collision/capt.hh:366-407, 425-506          point-cloud tree query
//   robots/<robot>.hh fkcc          link hierarchy: coarse sphere -> fine spheres; self pairs
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "robots_gen.h"

namespace
{
    constexpr int kRake = 8;

    // ---------------------------------------------------------------- trig (avx.hh:455-549)
    inline float vamp_sin(float x)
    {
        std::uint32_t sign_bit;
        std::memcpy(&sign_bit, &x, 4);
        sign_bit &= 0x80000000u;
        x = std::fabs(x);
        float y = x * 1.27323954473516f;
        int j = static_cast<int>(std::nearbyintf(y));  // _mm256_cvtps_epi32: round to nearest even
        j = (j + 1) & ~1;
        y = static_cast<float>(j);
        const std::uint32_t swap = (static_cast<std::uint32_t>(j) & 4u) << 29;
        const bool use_sin_poly = (j & 2) == 0;
        sign_bit ^= swap;
        x = x + y * -0.78515625f;
        x = x + y * -2.4187564849853515625e-4f;
        x = x + y * -3.77489497744594108e-8f;
        const float z = x * x;
        float yc = 2.443315711809948E-005f;
        yc = yc * z + -1.388731625493765E-003f;
        yc = yc * z + 4.166664568298827E-002f;
        yc = yc * z;
        yc = yc * z;
        yc = yc - z * 0.5f;
        yc = yc + 1.0f;
        float ys = -1.9515295891E-4f;
        ys = ys * z + 8.3321608736E-3f;
        ys = ys * z + -1.6666654611E-1f;
        ys = ys * z;
        ys = ys * x;
        ys = ys + x;
        float r = use_sin_poly ? ys : yc;
        std::uint32_t bits;
        std::memcpy(&bits, &r, 4);
        bits ^= sign_bit;
        std::memcpy(&r, &bits, 4);
        return r;
    }

    inline float vamp_cos(float x)  // interface.hh:451-459
    {
        constexpr float PI = 3.14159265359;
        const float v = x + static_cast<float>(PI / 2.);
        const float w = v - ((v >= PI) ? static_cast<float>(2 * PI) : 0.0f);
        return vamp_sin(w);
    }

    // ---------------------------------------------------------------- scene
    struct Sphere { float x, y, z, r, min_d; };
    struct Capsule { float x1, y1, z1, xv, yv, zv, r, rdv, min_d; };
    struct Cuboid { float x, y, z, a1x, a1y, a1z, a2x, a2y, a2z, a3x, a3y, a3z, h1, h2, h3, min_d; };

    struct Capt
    {
        int nlog2 = 0;
        std::vector<float> tests;
        std::vector<float> aabbs;  // [leaf][6]
        std::vector<std::uint32_t> aff_starts;
        std::vector<float> aff;  // [3][n_aff*8]
        std::size_t n_aff8 = 0;
        float top[6];
        float r_min, r_max, r_point;
    };

    // shapes.hh:250-312; xs, ys, zs are the RECIPROCAL scales (factory.hh:363-386)
    struct HeightField
    {
        float x, y, z, xs, ys, zs;
        std::size_t xd, yd, xd2, yd2;
        std::vector<float> data;
    };

    struct Env
    {
        std::vector<Sphere> spheres;
        std::vector<Capsule> capsules, zcapsules;
        std::vector<Cuboid> cuboids, zcuboids;
        std::vector<HeightField> hfs;
        std::vector<Capt> pcs;
        // collision/attachments.hh:13-56: pose relative to the end effector (row-major 3x4) + spheres in the attachment's frame
        bool has_att = false;
        float att_tf[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        std::vector<Sphere> att;
    };

    inline float dot3(float ax, float ay, float az, float bx, float by, float bz) { return ax * bx + ay * by + az * bz; }
    inline float clampf(float v, float lo, float hi) { return std::max(std::min(v, hi), lo); }  // math.hh:51-54

    // primitive tests: value < 0 (sign bit) <=> collision ---------------------------------------
    inline float t_sphere(const Sphere &o, float x, float y, float z, float r)  // sphere_sphere.hh:10-23
    {
        const float xs = o.x - x, ys = o.y - y, zs = o.z - z;
        const float sum = dot3(xs, ys, zs, xs, ys, zs);
        const float rs = o.r + r;
        return sum - rs * rs;
    }
    inline float t_capsule(const Capsule &c, float x, float y, float z, float r)  // sphere_capsule.hh:9-23
    {
        const float dot = dot3(x - c.x1, y - c.y1, z - c.z1, c.xv, c.yv, c.zv);
        const float cdf = std::min(std::max(dot * c.rdv, 0.0f), 1.0f);  // avx.hh clamp = min(max(v,lo),hi)
        const float px = c.x1 + c.xv * cdf, py = c.y1 + c.yv * cdf, pz = c.z1 + c.zv * cdf;
        const float xs = x - px, ys = y - py, zs = z - pz;
        const float sum = dot3(xs, ys, zs, xs, ys, zs);
        const float rs = r + c.r;
        return sum - rs * rs;
    }
    inline float t_zcapsule(const Capsule &c, float x, float y, float z, float r)  // sphere_capsule.hh:31-44
    {
        const float dot = (z - c.z1) * c.zv;
        const float cdf = std::min(std::max(dot * c.rdv, 0.0f), 1.0f);
        const float xs = x - c.x1, ys = y - c.y1, zs = z - (c.z1 + c.zv * cdf);
        const float sum = dot3(xs, ys, zs, xs, ys, zs);
        const float rs = r + c.r;
        return sum - rs * rs;
    }
    inline float t_cuboid(const Cuboid &c, float x, float y, float z, float rsq)  // sphere_cuboid.hh:9-27
    {
        const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
        const float a1 = std::max(std::fabs(dot3(c.a1x, c.a1y, c.a1z, xs, ys, zs)) - c.h1, 0.0f);
        const float a2 = std::max(std::fabs(dot3(c.a2x, c.a2y, c.a2z, xs, ys, zs)) - c.h2, 0.0f);
        const float a3 = std::max(std::fabs(dot3(c.a3x, c.a3y, c.a3z, xs, ys, zs)) - c.h3, 0.0f);
        return dot3(a1, a2, a3, a1, a2, a3) - rsq;
    }
    inline float t_zcuboid(const Cuboid &c, float x, float y, float z, float rsq)  // sphere_cuboid.hh:35-52
    {
        const float xs = x - c.x, ys = y - c.y, zs = z - c.z;
        const float a1 = std::max(std::fabs(c.a1x * xs + c.a1y * ys) - c.h1, 0.0f);
        const float a2 = std::max(std::fabs(c.a2x * xs + c.a2y * ys) - c.h2, 0.0f);
        const float a3 = std::max(std::fabs(zs) - c.h3, 0.0f);
        return dot3(a1, a2, a3, a1, a2, a3) - rsq;
    }

    inline float aabb_distsq(const float *bb, const float *p)  // capt.hh:47-54 (Volume::distsq_to)
    {
        float s = 0;
        for (int k = 0; k < 3; ++k)
        {
            const float d = p[k] - std::min(std::max(p[k], bb[k]), bb[3 + k]);
            s += d * d;
        }
        return s;
    }

    // capt.hh:366-407 (scalar `collides`; `collides_simd` :425-506 makes the same decision per lane)
    inline bool capt_collides(const Capt &t, float x, float y, float z, float r)
    {
        r = r + t.r_point;
        const float rsq = r * r;
        const float c[3] = {x, y, z};
        if (aabb_distsq(t.top, c) > rsq)
        {
            return false;
        }
        std::size_t idx = 0;
        for (int i = 0, k = 0; i < t.nlog2; ++i)
        {
            idx = 2 * idx + 1 + (c[k] >= t.tests[idx] ? 1 : 0);
            k = (k + 1) % 3;
        }
        const std::size_t zl = idx - t.tests.size();
        if (aabb_distsq(&t.aabbs[6 * zl], c) > rsq)
        {
            return false;
        }
        const std::size_t n8 = t.n_aff8;
        for (std::uint32_t i = t.aff_starts[zl]; i < t.aff_starts[zl + 1]; ++i)
        {
            for (int l = 0; l < 8; ++l)
            {
                const float dx = t.aff[(0 * n8 + i) * 8 + l] - x;
                const float dy = t.aff[(1 * n8 + i) * 8 + l] - y;
                const float dz = t.aff[(2 * n8 + i) * 8 + l] - z;
                if (dx * dx + dy * dy + dz * dz <= rsq)
                {
                    return true;
                }
            }
        }
        return false;
    }

    // ---------------------------------------------------------------- 8-lane environment test
    // validity.hh:46-150. Lanes = the 8 time samples of one rake; every break / early return is joint.
    struct Counters
    {
        double flops = 0;  // W_ref: flops of the vector ops the reference would execute, x8 lanes
    };

    template <typename Vec, typename Test>
    inline bool scan_class(const Vec &v, const float *ext, int lanes, int f_test, Counters *cnt, Test test)
    {
        for (const auto &o : v)
        {
            bool all_pruned = true;
            for (int l = 0; l < lanes; ++l)
            {
                if (o.min_d - ext[l] < 0.0f)
                {
                    all_pruned = false;
                }
            }
            if (cnt) cnt->flops += 1.0 * kRake;
            if (all_pruned)
            {
                break;
            }
            bool any = false;
            for (int l = 0; l < lanes; ++l)
            {
                if (std::signbit(test(o, l)))
                {
                    any = true;
                }
            }
            if (cnt) cnt->flops += static_cast<double>(f_test) * kRake;
            if (any)
            {
                return true;
            }
        }
        return false;
    }

    // sphere_heightfield.hh:9-30: value < 0 (sign bit) <=> collision. The cell index is clamped to [0, xd] x [0, yd] -- one past the
    // last column / row, exactly as the reference does; reads beyond the array (sphere past the far edge) are undefined there and
    // clamped to the last element here.
    inline float t_heightfield(const HeightField &a, float x, float y, float z, float r)
    {
        const float xo = a.x - x, yo = a.y - y;
        const float xs = std::floor(clampf(std::fmaf(a.xs, xo, static_cast<float>(a.xd2)), 0.F, static_cast<float>(a.xd)));
        const float ys = std::floor(clampf(std::fmaf(a.ys, yo, static_cast<float>(a.yd2)), 0.F, static_cast<float>(a.yd)));
        const float index = std::fmaf(ys, static_cast<float>(a.xd), xs);
        std::size_t i = static_cast<std::size_t>(std::lrintf(index));
        if (i >= a.data.size()) i = a.data.size() - 1;
        const float zhs = std::fmaf(a.zs, a.data[i], a.z);
        return (z - r) - zhs;
    }

    inline bool sphere_env(const Env &e, const float *sx, const float *sy, const float *sz, float sr, int lanes, Counters *cnt)
    {
        float ext[kRake];
        for (int l = 0; l < lanes; ++l)
        {
            ext[l] = std::sqrt(dot3(sx[l], sy[l], sz[l], sx[l], sy[l], sz[l])) + sr;  // reference: v*rsqrt(v), ~12 bit
        }
        if (cnt) cnt->flops += 7.0 * kRake;
        if (scan_class(e.spheres, ext, lanes, 11, cnt, [&](const Sphere &o, int l) { return t_sphere(o, sx[l], sy[l], sz[l], sr); }))
            return true;
        if (scan_class(e.capsules, ext, lanes, 26, cnt, [&](const Capsule &o, int l) { return t_capsule(o, sx[l], sy[l], sz[l], sr); }))
            return true;
        if (scan_class(e.zcapsules, ext, lanes, 18, cnt, [&](const Capsule &o, int l) { return t_zcapsule(o, sx[l], sy[l], sz[l], sr); }))
            return true;
        const float rsq = sr * sr;
        if (scan_class(e.cuboids, ext, lanes, 27, cnt, [&](const Cuboid &o, int l) { return t_cuboid(o, sx[l], sy[l], sz[l], rsq); }))
            return true;
        if (scan_class(e.zcuboids, ext, lanes, 22, cnt, [&](const Cuboid &o, int l) { return t_zcuboid(o, sx[l], sy[l], sz[l], rsq); }))
            return true;
        for (const auto &hf : e.hfs)  // validity.hh:132-138 (no pruning: min_distance is 0)
        {
            for (int l = 0; l < lanes; ++l)
            {
                if (t_heightfield(hf, sx[l], sy[l], sz[l], sr) < 0.0f)
                {
                    return true;
                }
            }
        }
        for (const auto &pc : e.pcs)
        {
            for (int l = 0; l < lanes; ++l)
            {
                if (capt_collides(pc, sx[l], sy[l], sz[l], sr))
                {
                    return true;
                }
            }
        }
        return false;
    }

    // ---------------------------------------------------------------- table-driven FK (float)
    struct Fr { float r[9]; float t[3]; };

    inline void fk_frames(const OrRobot &R, const float *q, Fr *F)
    {
        for (int f = 0; f < R.n_frames; ++f)
        {
            const OrFrame &d = R.frames[f];
            if (d.parent < 0)
            {
                const float id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
                std::memcpy(F[f].r, id, sizeof(id));
                F[f].t[0] = F[f].t[1] = F[f].t[2] = 0;
                continue;
            }
            const Fr &P = F[d.parent];
            float A[9], t[3];
            for (int i = 0; i < 3; ++i)
            {
                for (int j = 0; j < 3; ++j)
                {
                    float acc = 0;
                    for (int k = 0; k < 3; ++k)
                    {
                        const float c = static_cast<float>(d.R[3 * k + j]);
                        if (std::fabs(d.R[3 * k + j]) >= 1e-9) acc += P.r[3 * i + k] * c;
                    }
                    A[3 * i + j] = acc;
                }
                float acc = P.t[i];
                for (int k = 0; k < 3; ++k)
                {
                    if (std::fabs(d.t[k]) >= 1e-12) acc += P.r[3 * i + k] * static_cast<float>(d.t[k]);
                }
                t[i] = acc;
            }
            const float qv = q[d.q_index];
            int ai = 0;
            for (int k = 1; k < 3; ++k)
            {
                if (std::fabs(d.axis[k]) > std::fabs(d.axis[ai])) ai = k;
            }
            const float sg = d.axis[ai] > 0 ? 1.0f : -1.0f;
            Fr &O = F[f];
            if (d.type == 0)
            {
                const float s = sg * vamp_sin(qv), c = vamp_cos(qv);
                const int u = (ai + 1) % 3, v = (ai + 2) % 3;
                for (int i = 0; i < 3; ++i)
                {
                    const float cu = A[3 * i + u], cv = A[3 * i + v];
                    O.r[3 * i + ai] = A[3 * i + ai];
                    O.r[3 * i + u] = cu * c + cv * s;
                    O.r[3 * i + v] = cv * c - cu * s;
                    O.t[i] = t[i];
                }
            }
            else
            {
                for (int i = 0; i < 3; ++i)
                {
                    for (int j = 0; j < 3; ++j) O.r[3 * i + j] = A[3 * i + j];
                    O.t[i] = t[i] + A[3 * i + ai] * (sg * qv);
                }
            }
        }
    }

    inline void place(const Fr &F, const OrSphere &s, float &x, float &y, float &z)
    {
        const float lx = static_cast<float>(s.c[0]), ly = static_cast<float>(s.c[1]), lz = static_cast<float>(s.c[2]);
        x = F.r[0] * lx + F.r[1] * ly + F.r[2] * lz + F.t[0];
        y = F.r[3] * lx + F.r[4] * ly + F.r[5] * lz + F.t[1];
        z = F.r[6] * lx + F.r[7] * ly + F.r[8] * lz + F.t[2];
    }

    // Attachment::pose (attachments.hh:43-56) at Robot::eefk(q) = frame[ee_frame] * constant offset: n_tf = p_tf * tf, then
    // n_tf * sphere centre. (x, y, z)[i] receive the world positions of the attachment's spheres.
    inline void pose_attachment(const OrRobot &R, const Env &e, const Fr *F, float *x, float *y, float *z)
    {
        const Fr &P = F[R.ee_frame];
        Fr E, N;
        for (int i = 0; i < 3; ++i)
        {
            for (int j = 0; j < 3; ++j)
            {
                float acc = 0;
                for (int k = 0; k < 3; ++k)
                {
                    if (std::fabs(R.ee_R[3 * k + j]) >= 1e-9) acc += P.r[3 * i + k] * static_cast<float>(R.ee_R[3 * k + j]);
                }
                E.r[3 * i + j] = acc;
            }
            float acc = P.t[i];
            for (int k = 0; k < 3; ++k)
            {
                if (std::fabs(R.ee_t[k]) >= 1e-12) acc += P.r[3 * i + k] * static_cast<float>(R.ee_t[k]);
            }
            E.t[i] = acc;
        }
        const float *t = e.att_tf;
        for (int i = 0; i < 3; ++i)
        {
            for (int j = 0; j < 3; ++j)
            {
                N.r[3 * i + j] = E.r[3 * i] * t[j] + E.r[3 * i + 1] * t[4 + j] + E.r[3 * i + 2] * t[8 + j];
            }
            N.t[i] = E.r[3 * i] * t[3] + E.r[3 * i + 1] * t[7] + E.r[3 * i + 2] * t[11] + E.t[i];
        }
        for (std::size_t s = 0; s < e.att.size(); ++s)
        {
            const Sphere &a = e.att[s];
            x[s] = N.r[0] * a.x + N.r[1] * a.y + N.r[2] * a.z + N.t[0];
            y[s] = N.r[3] * a.x + N.r[4] * a.y + N.r[5] * a.z + N.t[1];
            z[s] = N.r[6] * a.x + N.r[7] * a.y + N.r[8] * a.z + N.t[2];
        }
    }

    constexpr int kMaxFrames = 16;
    constexpr int kMaxAtt = 256;
    constexpr int kMaxLinks = 48;
    constexpr int kMaxFine = 160;

    // Robot::fkcc<rake> on `lanes` configurations (robots/panda.hh:5226-10263): true = all lanes collision free.
    inline bool fkcc(const OrRobot &R, const Env &e, const float (*q)[16], int lanes, Counters *cnt, double fk_flops)
    {
        static thread_local float cx[kMaxLinks][kRake], cy[kMaxLinks][kRake], cz[kMaxLinks][kRake];
        static thread_local float fx[kMaxFine][kRake], fy[kMaxFine][kRake], fz[kMaxFine][kRake];
        static thread_local float ax[kMaxAtt][kRake], ay[kMaxAtt][kRake], az[kMaxAtt][kRake];
        const int n_att = e.has_att ? static_cast<int>(e.att.size()) : 0;
        Fr F[kMaxFrames];
        for (int l = 0; l < lanes; ++l)
        {
            fk_frames(R, q[l], F);
            if (n_att > 0)
            {
                float tx[kMaxAtt], ty[kMaxAtt], tz[kMaxAtt];
                pose_attachment(R, e, F, tx, ty, tz);
                for (int s = 0; s < n_att; ++s)
                {
                    ax[s][l] = tx[s];
                    ay[s][l] = ty[s];
                    az[s][l] = tz[s];
                }
            }
            for (int k = 0; k < R.n_links; ++k)
            {
                const OrLink &L = R.links[k];
                place(F[L.frame], L.coarse, cx[k][l], cy[k][l], cz[k][l]);
                for (int s = L.fine_begin; s < L.fine_end; ++s)
                {
                    place(F[L.frame], R.fine[s], fx[s][l], fy[s][l], fz[s][l]);
                }
            }
        }
        if (cnt) cnt->flops += fk_flops * kRake;
        // environment vs robot, links in the reference's order
        for (int k = 0; k < R.n_links; ++k)
        {
            const OrLink &L = R.links[k];
            if (sphere_env(e, cx[k], cy[k], cz[k], static_cast<float>(L.coarse.r), lanes, cnt))
            {
                for (int s = L.fine_begin; s < L.fine_end; ++s)
                {
                    if (sphere_env(e, fx[s], fy[s], fz[s], static_cast<float>(R.fine[s].r), lanes, cnt))
                    {
                        return false;
                    }
                }
            }
        }
        // self collisions (validity.hh:13-44)
        auto pair_hit = [&](const float *ax, const float *ay, const float *az, float ar, const float *bx, const float *by,
                            const float *bz, float br)
        {
            if (cnt) cnt->flops += 11.0 * kRake;
            for (int l = 0; l < lanes; ++l)
            {
                const float xs = ax[l] - bx[l], ys = ay[l] - by[l], zs = az[l] - bz[l];
                const float rs = ar + br;
                if (std::signbit(dot3(xs, ys, zs, xs, ys, zs) - rs * rs))
                {
                    return true;
                }
            }
            return false;
        };
        for (int p = 0; p < R.n_pairs; ++p)
        {
            const int a = R.pairs[p][0], b = R.pairs[p][1];
            const OrLink &A = R.links[a], &B = R.links[b];
            if (pair_hit(cx[a], cy[a], cz[a], static_cast<float>(A.coarse.r), cx[b], cy[b], cz[b], static_cast<float>(B.coarse.r)))
            {
                for (int i = A.fine_begin; i < A.fine_end; ++i)
                {
                    for (int j = B.fine_begin; j < B.fine_end; ++j)
                    {
                        if (pair_hit(fx[i], fy[i], fz[i], static_cast<float>(R.fine[i].r), fx[j], fy[j], fz[j],
                                     static_cast<float>(R.fine[j].r)))
                        {
                            return false;
                        }
                    }
                }
            }
        }
        if (n_att > 0)
        {
            // Robot::fkcc_attach (robots/panda.hh:15308-15466): attachment vs environment (validity.hh:251-268) ...
            for (int s = 0; s < n_att; ++s)
            {
                if (sphere_env(e, ax[s], ay[s], az[s], e.att[s].r, lanes, cnt))
                {
                    return false;
                }
            }
            // ... then vs the listed links: bounding sphere, then fine spheres (attachment_sphere_collision, validity.hh:270-293)
            auto att_hit = [&](const float *sx, const float *sy, const float *sz, float sr)
            {
                for (int s = 0; s < n_att; ++s)
                {
                    if (pair_hit(sx, sy, sz, sr, ax[s], ay[s], az[s], e.att[s].r))
                    {
                        return true;
                    }
                }
                return false;
            };
            for (int k = 0; k < R.n_att_links; ++k)
            {
                const int li = R.att_links[k];
                const OrLink &L = R.links[li];
                if (att_hit(cx[li], cy[li], cz[li], static_cast<float>(L.coarse.r)))
                {
                    for (int s = L.fine_begin; s < L.fine_end; ++s)
                    {
                        if (att_hit(fx[s], fy[s], fz[s], static_cast<float>(R.fine[s].r)))
                        {
                            return false;
                        }
                    }
                }
            }
        }
        return true;
    }

    // avx.hh:442-452 hsum of the squared lanes (pad lanes are zero)
    inline float l2_norm(const float *v, int dim)
    {
        float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        float total = 0;
        for (int base = 0; base < dim; base += 8)  // one __m256 per 8 scalars; rows summed first (unpack::sum_)
        {
            for (int l = 0; l < 8 && base + l < dim; ++l)
            {
                s[l] += v[base + l] * v[base + l];
            }
        }
        const float s1[4] = {s[4] + s[0], s[5] + s[1], s[6] + s[2], s[7] + s[3]};
        const float s2a = s1[0] + s1[2];
        const float s2b = s1[1] + s1[3];
        total = s2a + s2b;
        return std::sqrt(total);
    }

    // validate.hh:23-77
    inline bool validate_motion(const OrRobot &R, const Env &e, const float *a, const float *b, int resolution, Counters *cnt,
                                double fk_flops)
    {
        const int d = R.dim;
        float vec[16], block[kRake][16];
        for (int j = 0; j < d; ++j)
        {
            vec[j] = b[j] - a[j];
        }
        const float dist = l2_norm(vec, d);
        for (int l = 0; l < kRake; ++l)
        {
            const float pct = static_cast<float>(l + 1) / static_cast<float>(kRake);
            for (int j = 0; j < d; ++j)
            {
                block[l][j] = a[j] + vec[j] * pct;
            }
        }
        const std::size_t n = static_cast<std::size_t>(
            std::max(std::ceil(dist / static_cast<float>(kRake) * static_cast<float>(resolution)), 1.0f));
        bool valid = fkcc(R, e, block, kRake, cnt, fk_flops);
        if (!valid || n == 1)
        {
            return valid;
        }
        float back[16];
        for (int j = 0; j < d; ++j)
        {
            back[j] = vec[j] / static_cast<float>(kRake * n);
        }
        for (std::size_t i = 1; i < n; ++i)
        {
            for (int l = 0; l < kRake; ++l)
            {
                for (int j = 0; j < d; ++j)
                {
                    block[l][j] = block[l][j] - back[j];
                }
            }
            if (!fkcc(R, e, block, kRake, cnt, fk_flops))
            {
                return false;
            }
        }
        return true;
    }

    // ------------------------------------------------------------ instrumentation: exhaustive clearance
    // Signed clearance (metres, double) of one robot sphere against every obstacle, no pruning, no hierarchy.
    inline double clearance_env(const Env &e, double x, double y, double z, double r)
    {
        double best = std::numeric_limits<double>::infinity();
        for (const auto &o : e.spheres)
        {
            const double dx = o.x - x, dy = o.y - y, dz = o.z - z;
            best = std::min(best, std::sqrt(dx * dx + dy * dy + dz * dz) - (r + o.r));
        }
        auto caps = [&](const Capsule &c)
        {
            const double dot = (x - c.x1) * c.xv + (y - c.y1) * c.yv + (z - c.z1) * c.zv;
            const double cdf = std::min(std::max(dot * static_cast<double>(c.rdv), 0.0), 1.0);
            const double dx = x - (c.x1 + c.xv * cdf), dy = y - (c.y1 + c.yv * cdf), dz = z - (c.z1 + c.zv * cdf);
            best = std::min(best, std::sqrt(dx * dx + dy * dy + dz * dz) - (r + c.r));
        };
        for (const auto &c : e.capsules) caps(c);
        for (const auto &c : e.zcapsules) caps(c);
        auto cub = [&](const Cuboid &c)
        {
            const double xs = x - c.x, ys = y - c.y, zs = z - c.z;
            const double a1 = std::max(std::fabs(c.a1x * xs + c.a1y * ys + c.a1z * zs) - c.h1, 0.0);
            const double a2 = std::max(std::fabs(c.a2x * xs + c.a2y * ys + c.a2z * zs) - c.h2, 0.0);
            const double a3 = std::max(std::fabs(c.a3x * xs + c.a3y * ys + c.a3z * zs) - c.h3, 0.0);
            best = std::min(best, std::sqrt(a1 * a1 + a2 * a2 + a3 * a3) - r);
        };
        for (const auto &c : e.cuboids) cub(c);
        for (const auto &c : e.zcuboids) cub(c);
        for (const auto &hf : e.hfs) best = std::min(best, static_cast<double>(t_heightfield(hf, static_cast<float>(x), static_cast<float>(y), static_cast<float>(z), static_cast<float>(r))));
        for (const auto &t : e.pcs)
        {
            const std::size_t n8 = t.n_aff8;
            // exhaustive over all stored affordance points (superset of the cloud; +inf pads never win)
            for (std::size_t i = 0; i < n8 * 8; ++i)
            {
                const double dx = t.aff[0 * n8 * 8 + i] - x, dy = t.aff[1 * n8 * 8 + i] - y, dz = t.aff[2 * n8 * 8 + i] - z;
                const double dd = std::sqrt(dx * dx + dy * dy + dz * dz) - (r + t.r_point);
                if (dd < best) best = dd;
            }
        }
        return best;
    }

    inline double state_clearance(const OrRobot &R, const Env &e, const float *q)
    {
        Fr F[kMaxFrames];
        fk_frames(R, q, F);
        static thread_local float px[kMaxFine], py[kMaxFine], pz[kMaxFine];
        double best = std::numeric_limits<double>::infinity();
        for (int k = 0; k < R.n_links; ++k)
        {
            const OrLink &L = R.links[k];
            for (int s = L.fine_begin; s < L.fine_end; ++s)
            {
                place(F[L.frame], R.fine[s], px[s], py[s], pz[s]);
                best = std::min(best, clearance_env(e, px[s], py[s], pz[s], static_cast<float>(R.fine[s].r)));
            }
        }
        for (int p = 0; p < R.n_pairs; ++p)
        {
            const OrLink &A = R.links[R.pairs[p][0]], &B = R.links[R.pairs[p][1]];
            for (int i = A.fine_begin; i < A.fine_end; ++i)
            {
                for (int j = B.fine_begin; j < B.fine_end; ++j)
                {
                    const double dx = px[i] - px[j], dy = py[i] - py[j], dz = pz[i] - pz[j];
                    best = std::min(best, std::sqrt(dx * dx + dy * dy + dz * dz) -
                                              (static_cast<double>(static_cast<float>(R.fine[i].r)) + static_cast<float>(R.fine[j].r)));
                }
            }
        }
        if (e.has_att && !e.att.empty())
        {
            float tx[kMaxAtt], ty[kMaxAtt], tz[kMaxAtt];
            pose_attachment(R, e, F, tx, ty, tz);
            for (std::size_t s = 0; s < e.att.size(); ++s)
            {
                best = std::min(best, clearance_env(e, tx[s], ty[s], tz[s], e.att[s].r));
                for (int k = 0; k < R.n_att_links; ++k)
                {
                    const OrLink &L = R.links[R.att_links[k]];
                    for (int i = L.fine_begin; i < L.fine_end; ++i)
                    {
                        const double dx = px[i] - tx[s], dy = py[i] - ty[s], dz = pz[i] - tz[s];
                        best = std::min(best, std::sqrt(dx * dx + dy * dy + dz * dz) - (static_cast<double>(static_cast<float>(R.fine[i].r)) + e.att[s].r));
                    }
                }
            }
        }
        return best;
    }

    inline const OrRobot *robot_by_id(int id)
    {
        switch (id)
        {
            case 0: return &panda_robot;
            case 1: return &fetch_robot;
            case 2: return &ur5_robot;
            case 3: return &baxter_robot;
            default: return nullptr;
        }
    }

    // flops of the reference's FK per state: counted muls+adds of the traced program + 32 per trig call
    // (Panda 614 mul + 533 add + 14 trig; Fetch 516 mul + 510 add + 14 trig)
    inline double fk_flops_of(int id) { return id == 0 ? (614 + 533 + 14 * 32) : (516 + 510 + 14 * 32); }

    template <typename F>
    void parallel_for(std::int64_t m, int nthreads, F body)
    {
        if (nthreads <= 1)
        {
            body(0, m);
            return;
        }
        std::atomic<std::int64_t> next{0};
        const std::int64_t chunk = 64;
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t)
        {
            pool.emplace_back(
                [&]()
                {
                    for (;;)
                    {
                        const std::int64_t b = next.fetch_add(chunk);
                        if (b >= m) break;
                        body(b, std::min(b + chunk, m));
                    }
                });
        }
        for (auto &t : pool) t.join();
    }
}  // namespace

extern "C"
{
    void *or_env_create() { return new Env(); }
    void or_env_destroy(void *e) { delete static_cast<Env *>(e); }

    void or_env_add_sphere(void *e, float x, float y, float z, float r)
    {
        Sphere s{x, y, z, r, 0};
        s.min_d = std::sqrt(x * x + y * y + z * z) - r;  // shapes.hh:236-239
        static_cast<Env *>(e)->spheres.push_back(s);
    }

    // tf: row-major 3x4 relative to the end effector; xyzr[n][4]; n == 0 removes the attachment (at most kMaxAtt spheres)
    int or_env_set_attachment(void *e, const float *tf, const float *xyzr, std::int64_t n)
    {
        auto &env = *static_cast<Env *>(e);
        env.att.clear();
        env.has_att = false;
        if (n <= 0) return 0;
        if (n > kMaxAtt) return -1;
        std::memcpy(env.att_tf, tf, sizeof(env.att_tf));
        for (std::int64_t i = 0; i < n; ++i) env.att.push_back(Sphere{xyzr[4 * i], xyzr[4 * i + 1], xyzr[4 * i + 2], xyzr[4 * i + 3], 0});
        env.has_att = true;
        return 0;
    }

    void or_env_add_cuboid(void *e, const float *c, int z_aligned)
    {
        Cuboid s{c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9], c[10], c[11], c[12], c[13], c[14], 0};
        // shapes.hh:52-67
        const float d1 = dot3(-s.x, -s.y, -s.z, s.a1x, s.a1y, s.a1z);
        const float d2 = dot3(-s.x, -s.y, -s.z, s.a2x, s.a2y, s.a2z);
        const float d3 = dot3(-s.x, -s.y, -s.z, s.a3x, s.a3y, s.a3z);
        const float v1 = clampf(d1, -s.h1, s.h1), v2 = clampf(d2, -s.h2, s.h2), v3 = clampf(d3, -s.h3, s.h3);
        const float xn = s.x + s.a1x * v1 + s.a2x * v2 + s.a3x * v3;
        const float yn = s.y + s.a1y * v1 + s.a2y * v2 + s.a3y * v3;
        const float zn = s.z + s.a1z * v1 + s.a2z * v2 + s.a3z * v3;
        s.min_d = std::sqrt(xn * xn + yn * yn + zn * zn);
        auto &env = *static_cast<Env *>(e);
        (z_aligned ? env.zcuboids : env.cuboids).push_back(s);
    }

    void or_env_add_capsule(void *e, const float *c, int z_aligned)
    {
        Capsule s{c[0], c[1], c[2], c[3], c[4], c[5], c[6], 0, 0};
        s.rdv = static_cast<float>(1.0 / dot3(s.xv, s.yv, s.zv, s.xv, s.yv, s.zv));  // factory.hh:110-126
        // shapes.hh:165-189
        const float dot = clampf(dot3(-s.x1, -s.y1, -s.z1, s.xv, s.yv, s.zv) * s.rdv, 0.F, 1.F);
        const float xp = s.x1 + s.xv * dot, yp = s.y1 + s.yv * dot, zp = s.z1 + s.zv * dot;
        float xo = -xp, yo = -yp, zo = -zp;
        const float ol = std::sqrt(dot3(xo, yo, zo, xo, yo, zo));
        xo = xo / ol;
        yo = yo / ol;
        zo = zo / ol;
        const float ro = clampf(ol, 0.F, s.r);
        const float xn = xp + ro * xo, yn = yp + ro * yo, zn = zp + ro * zo;
        s.min_d = std::sqrt(xn * xn + yn * yn + zn * zn);
        auto &env = *static_cast<Env *>(e);
        (z_aligned ? env.zcapsules : env.capsules).push_back(s);
    }

    // factory::heightfield::flat (factory.hh:363-386): centre, scale (stored as reciprocals), cells, row-major heights
    void or_env_add_heightfield(void *e, float cx, float cy, float cz, float sx, float sy, float sz, std::int64_t nx, std::int64_t ny, const float *data)
    {
        HeightField h;
        h.x = cx; h.y = cy; h.z = cz;
        h.xs = 1.F / sx; h.ys = 1.F / sy; h.zs = 1.F / sz;
        h.xd = static_cast<std::size_t>(nx); h.yd = static_cast<std::size_t>(ny);
        h.xd2 = h.xd / 2; h.yd2 = h.yd / 2;
        h.data.assign(data, data + nx * ny);
        static_cast<Env *>(e)->hfs.push_back(std::move(h));
    }

    // attach a point-cloud tree given as flat arrays (layout of refv_capt_export / b200mp_env_add_capt)
    void or_env_add_capt(void *e, int nlog2, const float *tests, const float *aabbs, const std::uint32_t *aff_starts,
                         const float *aff, std::int64_t n_aff8, const float *top, float r_min, float r_max, float r_point)
    {
        Capt t;
        t.nlog2 = nlog2;
        const std::size_t nt = (std::size_t(1) << nlog2) - 1, nl = std::size_t(1) << nlog2;
        t.tests.assign(tests, tests + nt);
        t.aabbs.assign(aabbs, aabbs + 6 * nl);
        t.aff_starts.assign(aff_starts, aff_starts + nl + 1);
        t.aff.assign(aff, aff + 3 * n_aff8 * 8);
        t.n_aff8 = static_cast<std::size_t>(n_aff8);
        std::memcpy(t.top, top, sizeof(t.top));
        t.r_min = r_min;
        t.r_max = r_max;
        t.r_point = r_point;
        static_cast<Env *>(e)->pcs.push_back(std::move(t));
    }

    void or_env_sort(void *e)  // environment.hh:43-69 (std::sort on min_distance)
    {
        auto &env = *static_cast<Env *>(e);
        auto by = [](const auto &a, const auto &b) { return a.min_d < b.min_d; };
        std::sort(env.spheres.begin(), env.spheres.end(), by);
        std::sort(env.capsules.begin(), env.capsules.end(), by);
        std::sort(env.zcapsules.begin(), env.zcapsules.end(), by);
        std::sort(env.cuboids.begin(), env.cuboids.end(), by);
        std::sort(env.zcuboids.begin(), env.zcuboids.end(), by);
    }

    std::int64_t or_env_min_distances(void *e, int kind, float *out)
    {
        auto &env = *static_cast<Env *>(e);
        auto dump = [&](const auto &v) -> std::int64_t
        {
            if (out) for (std::size_t i = 0; i < v.size(); ++i) out[i] = v[i].min_d;
            return static_cast<std::int64_t>(v.size());
        };
        switch (kind)
        {
            case 0: return dump(env.spheres);
            case 1: return dump(env.capsules);
            case 2: return dump(env.zcapsules);
            case 3: return dump(env.cuboids);
            case 4: return dump(env.zcuboids);
            default: return -1;
        }
    }

    // valid[m]; flops[m] (nullable) = W_ref per edge. resolution 0 -> robot default.
    int or_validate_motions(void *env, int robot, const float *q0, const float *q1, std::int64_t m, int resolution,
                            std::uint8_t *valid, double *flops, int nthreads)
    {
        const OrRobot *R = robot_by_id(robot);
        if (!R || !env) return -1;
        if (resolution <= 0) resolution = R->resolution;
        const Env &e = *static_cast<Env *>(env);
        const double fkf = fk_flops_of(robot);
        parallel_for(m, nthreads,
                     [&](std::int64_t b, std::int64_t en)
                     {
                         for (std::int64_t i = b; i < en; ++i)
                         {
                             Counters c;
                             valid[i] = validate_motion(*R, e, q0 + i * R->dim, q1 + i * R->dim, resolution, flops ? &c : nullptr, fkf) ? 1 : 0;
                             if (flops) flops[i] = c.flops;
                         }
                     });
        return 0;
    }

    // VampStateValidityChecker::isValid = validate_motion<Robot, rake, 1>(q, q)
    int or_validate_states(void *env, int robot, const float *q, std::int64_t m, std::uint8_t *valid, int nthreads)
    {
        return or_validate_motions(env, robot, q, q, m, 1, valid, nullptr, nthreads);
    }

    // Minimum signed clearance (double, metres) over every state the reference would test on the edge,
    // every fine sphere, every obstacle and every allowed self pair -- no pruning, no hierarchy, no early exit.
    int or_edge_min_clearance(void *env, int robot, const float *q0, const float *q1, std::int64_t m, int resolution,
                              double *out, int nthreads)
    {
        const OrRobot *R = robot_by_id(robot);
        if (!R || !env) return -1;
        if (resolution <= 0) resolution = R->resolution;
        const Env &e = *static_cast<Env *>(env);
        const int d = R->dim;
        parallel_for(m, nthreads,
                     [&](std::int64_t b, std::int64_t en)
                     {
                         for (std::int64_t i = b; i < en; ++i)
                         {
                             const float *a = q0 + i * d, *bb = q1 + i * d;
                             float vec[16], back[16], blk[kRake][16];
                             for (int j = 0; j < d; ++j) vec[j] = bb[j] - a[j];
                             const float dist = l2_norm(vec, d);
                             const std::size_t n = static_cast<std::size_t>(
                                 std::max(std::ceil(dist / static_cast<float>(kRake) * static_cast<float>(resolution)), 1.0f));
                             for (int l = 0; l < kRake; ++l)
                                 for (int j = 0; j < d; ++j)
                                     blk[l][j] = a[j] + vec[j] * (static_cast<float>(l + 1) / static_cast<float>(kRake));
                             for (int j = 0; j < d; ++j) back[j] = vec[j] / static_cast<float>(kRake * n);
                             double best = std::numeric_limits<double>::infinity();
                             for (std::size_t r = 0; r < n; ++r)
                             {
                                 if (r > 0)
                                     for (int l = 0; l < kRake; ++l)
                                         for (int j = 0; j < d; ++j) blk[l][j] -= back[j];
                                 for (int l = 0; l < kRake; ++l) best = std::min(best, state_clearance(*R, e, blk[l]));
                             }
                             out[i] = best;
                         }
                     });
        return 0;
    }

    // fine spheres in the reference's sphere index order: out[m][n_fine][4]
    int or_sphere_fk(int robot, const float *q, std::int64_t m, float *out)
    {
        const OrRobot *R = robot_by_id(robot);
        if (!R) return -1;
        // tables are grouped by link; fine_id maps back to the reference's sphere numbering
        Fr F[kMaxFrames];
        for (std::int64_t i = 0; i < m; ++i)
        {
            fk_frames(*R, q + i * R->dim, F);
            for (int k = 0; k < R->n_links; ++k)
            {
                const OrLink &L = R->links[k];
                for (int s = L.fine_begin; s < L.fine_end; ++s)
                {
                    float *o = out + (i * R->n_fine + R->fine_id[s]) * 4;
                    place(F[L.frame], R->fine[s], o[0], o[1], o[2]);
                    o[3] = static_cast<float>(R->fine[s].r);
                }
            }
        }
        return 0;
    }

    // Robot::eefk (robots/panda.hh:15472-15552): out[m][12] = row-major 3x4 pose of the end effector (identity attachment posed there)
    int or_eefk(int robot, const float *q, std::int64_t m, float *out)
    {
        const OrRobot *R = robot_by_id(robot);
        if (!R) return -1;
        Env e;
        e.has_att = true;
        e.att = {Sphere{0, 0, 0, 0, 0}, Sphere{1, 0, 0, 0, 0}, Sphere{0, 1, 0, 0, 0}, Sphere{0, 0, 1, 0, 0}};
        Fr F[kMaxFrames];
        for (std::int64_t i = 0; i < m; ++i)
        {
            fk_frames(*R, q + i * R->dim, F);
            float x[4], y[4], z[4];
            pose_attachment(*R, e, F, x, y, z);
            float *o = out + i * 12;
            for (int c = 0; c < 3; ++c)
            {
                o[c] = x[1 + c] - x[0];
                o[4 + c] = y[1 + c] - y[0];
                o[8 + c] = z[1 + c] - z[0];
            }
            o[3] = x[0];
            o[7] = y[0];
            o[11] = z[0];
        }
        return 0;
    }

    int or_robot_info(int robot, int *dim, int *n_fine, int *n_links, int *resolution, float *lo, float *hi)
    {
        const OrRobot *R = robot_by_id(robot);
        if (!R) return -1;
        *dim = R->dim;
        *n_fine = R->n_fine;
        *n_links = R->n_links;
        *resolution = R->resolution;
        for (int j = 0; j < R->dim; ++j)
        {
            lo[j] = R->lo[j];
            hi[j] = R->hi[j];
        }
        return 0;
    }

    void or_trig(const float *x, std::int64_t n, float *s, float *c)
    {
        for (std::int64_t i = 0; i < n; ++i)
        {
            s[i] = vamp_sin(x[i]);
            c[i] = vamp_cos(x[i]);
        }
    }
}
