// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY.
// C wrapper around the three reference translation units that compile stand-alone
// (ohao/render/rt/{env_cdf,sobol_generator,owen_scramble}.cpp).  Built by
// `make ref` from the sources where they lie under $(REF); output goes to oracle/_ref/ only.
// Used to pin the oracle's restatement (tests/test_oracle_ref.py) and as generator of
// tests/golden/*.npz fixtures (tools/gen_golden_fixtures.py).
#include "render/rt/env_cdf.hpp"
#include "render/rt/sobol_generator.hpp"
#include "render/rt/owen_scramble.hpp"
#include <cstring>
extern "C" {
void ref_env_cdf(const float* rgba, int w, int h, float* marg, float* cond, float* integral) {
    ohao::EnvCDF cdf; cdf.build(rgba, w, h);
    std::memcpy(marg, cdf.marginalCDF().data(), sizeof(float) * cdf.marginalCDF().size());
    std::memcpy(cond, cdf.conditionalCDF().data(), sizeof(float) * cdf.conditionalCDF().size());
    *integral = cdf.integral();
}
float ref_sobol_sample1d(uint32_t index, uint32_t dim) { return ohao::SobolGenerator::sample1D(index, dim); }
const uint32_t* ref_sobol_dirs(void) { return ohao::SobolGenerator::directionNumbers(); }
uint32_t ref_owen(uint32_t v, uint32_t seed) { return ohao::owenScramble(v, seed); }
}
