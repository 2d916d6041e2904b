// rng_skip.h -- advance a generator exactly as ONE draw of a FRESH std::normal_distribution<double> would (the reference makes a new distribution
// object for every sample, src/random.cpp:69-73, so libstdc++'s saved second value is never used).  libstdc++ draws by the polar method
// (bits/random.tcc, normal_distribution::operator()): pairs of canonical doubles until they fall inside the unit circle; the value needs a log
// and a sqrt on top, the generator state does not.  A sharded rank uses this for the particles of OTHER ranks on scans that will be matched: their
// predicted poses are replaced by the gathered match results, only the random stream has to stay aligned.
#pragma once

#include <limits>
#include <random>

namespace lama_b200 {

template <typename Gen>
inline void rng_skip_normal(Gen& gen)
{
    double x, y, r2;
    do {
        x  = 2.0 * std::generate_canonical<double, std::numeric_limits<double>::digits>(gen) - 1.0;
        y  = 2.0 * std::generate_canonical<double, std::numeric_limits<double>::digits>(gen) - 1.0;
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0.0);
}

}  // namespace lama_b200
