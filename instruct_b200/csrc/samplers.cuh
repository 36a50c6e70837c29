// samplers.cuh -- gamma / normal / beta draws on the Philox stream, host + device.
//
// They replace rgamma1 / rgamma2 / rexp / rbeta / rstd_normal of the reference
// (random.c:121-307).  The target distributions are identical; the algorithms
// (Marsaglia-Tsang 2000 squeeze for the gamma, with the U^(1/a) boost for shape < 1, and
// Box-Muller for the normal) and the streams are not -- BASELINE.json's north_star makes
// posterior agreement, not stream agreement, the acceptance test.
#pragma once
#include <math.h>
#include "philox.cuh"

namespace ig {

IG_HD double draw_normal(Stream &st)
{
	const double u1 = st.uniform(), u2 = st.uniform();
	return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
}

IG_HD double draw_gamma(Stream &st, double a)
{
	double boost = 1.0;
	if (a < 1.0) {
		boost = pow(st.uniform(), 1.0 / a);
		a += 1.0;
	}
	const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
	for (int it = 0; it < 1000; ++it) {
		const double x = draw_normal(st);
		double v = 1.0 + c * x;
		if (v <= 0.0) continue;
		v = v * v * v;
		const double u = st.uniform();
		if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) {
			// May underflow to exactly 0 for tiny shapes, like the reference's pow(x, 1/alpha)
			// (random.c:187): update_alpha's behaviour at q == 0 is part of the chain (see post_sweep).
			return d * v * boost;
		}
	}
	return d;
}

IG_HD double draw_beta(Stream &st, double a, double b)
{
	const double x = draw_gamma(st, a);
	return x / (x + draw_gamma(st, b));
}

}  // namespace ig
