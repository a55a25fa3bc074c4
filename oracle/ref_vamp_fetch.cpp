// TEST INFRASTRUCTURE ONLY -- see ref_vamp_impl.hh. Instantiates the reference for vamp::robots::Fetch.
#include "ref_vamp_impl.hh"
#include <vamp/robots/fetch.hh>

using Robot = vamp::robots::Fetch;

extern "C"
{
    int refv_fetch_dimension() { return static_cast<int>(Robot::dimension); }
    int refv_fetch_n_spheres() { return static_cast<int>(Robot::n_spheres); }
    int refv_fetch_resolution() { return static_cast<int>(Robot::resolution); }

    // lo[d], hi[d]: the bounds the OMPL bridge derives by scaling 0 and 1 (src/ompl/vamp/VampStateSpace.h:10-36)
    void refv_fetch_bounds(float *lo, float *hi)
    {
        std::array<float, Robot::Configuration::num_scalars> z{}, o{};
        o.fill(1.0F);
        Robot::Configuration a(z), b(o);
        Robot::scale_configuration(a);
        Robot::scale_configuration(b);
        const auto za = a.to_array();
        const auto oa = b.to_array();
        for (std::size_t i = 0; i < Robot::dimension; ++i)
        {
            lo[i] = za[i];
            hi[i] = oa[i];
        }
    }

    int refv_fetch_validate_motions(
        void *env, const float *q0, const float *q1, std::int64_t m, int resolution, std::uint8_t *out, int nthreads)
    {
        return refv::validate_motions<Robot>(static_cast<refv::Env *>(env), q0, q1, m, resolution, out, nthreads);
    }

    int refv_fetch_sphere_fk(const float *q, std::int64_t m, float *out) { return refv::sphere_fk<Robot>(q, m, out); }

    // out[m][12] = row-major 3x4 [R | t] of Robot::eefk (the frame fkcc_attach poses attachments in)
    int refv_fetch_eefk(const float *q, std::int64_t m, float *out) { return refv::eefk<Robot>(q, m, out); }
}
