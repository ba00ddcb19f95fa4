// hipace.dt = adaptive: the host arithmetic of src/utils/AdaptiveTimeStep.cpp (CalculateFromMinUz
// :143-233 incl. the look-ahead over the numprocs ranks of a pipeline :225-251, CalculateFromDensity
// :315-369).  Pure C++, exposed through hpb_adaptive_dt_* for the CPU tests.
#pragma once
#include <math.h>
#include <float.h>
#include <functional>

#include "../../include/hpb200.h"      // hpb_adaptive_par

// ts[4 * ib + {0,1,2,3}] = min uz / c, sum w, sum w uz / c, sum w uz^2 / c^2 of beam ib (:108-141).
// rho_at(z) = MultiPlasma::maxChargeDensity(z); t_now = the physical time of the step that just ended.
// Returns false on the reference's assertion failures.
inline bool adaptive_dt_from_min_uz(const hpb_adaptive_par &p, int nbeams, const double *ts,
                                    const double *charge, const double *mass,
                                    const std::function<double(double)> &rho_at, double t_now, double dt_in,
                                    double &dt_out, double &min_uz_mq)
{
    // the new dt is used numprocs time steps later: re-evaluate the betatron frequency at the predicted
    // times (hipace.adaptive_predict_step, default on)
    const int niter = (p.predict_step && p.numprocs > 1) ? p.numprocs : 1;
    double dt_min = HUGE_VAL, mq_min = DBL_MAX;
    for (int ib = 0; ib < nbeams; ++ib) {
        double new_dt = dt_in;
        if (charge[ib] != 0.) {
            const double mcr = mass[ib] / charge[ib];
            const double *t = ts + 4 * ib;
            if (t[1] == 0.) return false;                  // "The sum of all weights is 0!"
            const double mean = t[2] / t[1];
            const double sigma = sqrt(fabs(t[3] / t[1] - mean * mean));
            double chosen = fmin(fmax(mean - 4. * sigma, t[0]), 1.e30);
            chosen = fmax(chosen, p.threshold_uz);
            mq_min = fmin(mq_min, fabs(chosen * mcr));
            double new_time = t_now, min_uz = chosen;
            for (int it = 0; it < niter; ++it) {
                const double rho = rho_at(p.c * new_time);
                if (!(rho > 0.)) return false;             // "A >0 plasma density must be specified"
                min_uz = fmax(min_uz, 0.001 * p.threshold_uz);
                const double omega_b = sqrt(rho / (2. * fabs(min_uz * mcr) * p.ep0));
                const double cand = 2. * 3.14159265358979323846 / omega_b / p.nt_per_betatron;
                new_time += cand;
                if (min_uz > p.threshold_uz) new_dt = cand;
            }
        }
        dt_min = fmin(dt_min, new_dt);
    }
    min_uz_mq = mq_min;
    dt_out = fmin(dt_min, p.dt_max);
    return true;
}

// rho_at(z): MultiPlasma::maxChargeDensity
inline double adaptive_dt_from_density(const hpb_adaptive_par &p, double min_uz_mq, double t, double dt,
                                       const std::function<double(double)> &rho_at)
{
    if (!p.control_phase) return dt;
    const double dt_sub = dt / p.phase_substeps;
    double pa = 0., pa0 = 0.;
    const double omgb0 = sqrt(rho_at(p.c * t) / (2. * min_uz_mq * p.ep0));
    for (int i = 0; i < p.phase_substeps; ++i) {
        const double omgb = sqrt(rho_at(p.c * (t + i * dt_sub)) / (2. * min_uz_mq * p.ep0));
        pa += omgb * dt_sub;
        pa0 += omgb0 * dt_sub;
        if (fabs(pa - pa0) > 2. * 3.14159265358979323846 * p.phase_tolerance / p.nt_per_betatron) return i * dt_sub;
    }
    return dt;
}
