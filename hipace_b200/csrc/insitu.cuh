// In-situ beam diagnostics: the per-particle terms of BeamParticleContainer::InSituComputeDiags
// (src/particles/beam/BeamParticleContainer.cpp:476-557), host + device, and the NumPy-structured
// file format of src/utils/InsituUtil.H (JSON dtype header + raw little-endian records) that the
// reference's tools/read_insitu_diagnostics.py reads.
#pragma once
#include "shapes.cuh"

constexpr int kInsituNReal = 22;      // m_insitu_nrp; slot 22 of a record holds Np

// PlasmaParticleContainer::InSituComputeDiags (src/particles/plasma/PlasmaParticleContainer.cpp:443-526):
// out[15] = w', w' x, w' x^2, w' y, w' y^2, w' ux, w' ux^2, w' uy, w' uy^2, w' uz, w' uz^2, w' gamma,
// w' gamma^2, w (gamma - 1), 1 with gamma = (1 + ux^2 + uy^2 + psi^2) / (2 psi), uz = gamma - psi,
// w' = w gamma / psi
constexpr int kPlasmaInsituNReal = 14;
HPB_HD bool insitu_plasma_terms(bool valid, double x, double y, double ux_c, double uy_c, double psi,
                                double w0, double clight_inv, double radius_sq, double out[15])
{
    const double ux = ux_c * clight_inv, uy = uy_c * clight_inv;
    if (!valid || x * x + y * y > radius_sq) return false;
    const double gamma = (1.0 + ux * ux + uy * uy + psi * psi) / (2.0 * psi);
    const double uz = gamma - psi;
    const double w = w0 * gamma / psi;
    out[0] = w;
    out[1] = w * x;   out[2] = w * x * x;
    out[3] = w * y;   out[4] = w * y * y;
    out[5] = w * ux;  out[6] = w * ux * ux;
    out[7] = w * uy;  out[8] = w * uy * uy;
    out[9] = w * uz;  out[10] = w * uz * uz;
    out[11] = w * gamma; out[12] = w * gamma * gamma;
    out[13] = w0 * (gamma - 1.0);
    out[14] = 1.;
    return true;
}

// Fields::InSituComputeDiags (src/fields/Fields.cpp:1289-1347), one cell: Ex^2, Ey^2, Ez^2, Bx^2, By^2,
// Bz^2, ExmBy^2, EypBx^2, jz_beam, Ez jz_beam
HPB_HD void insitu_field_terms(double exmby, double eypbx, double ez, double bx, double by, double bz,
                               double jzb, double clight, double out[10])
{
    const double ex = exmby + by * clight, ey = eypbx - bx * clight;
    out[0] = ex * ex; out[1] = ey * ey; out[2] = ez * ez;
    out[3] = bx * bx; out[4] = by * by; out[5] = bz * bz;
    out[6] = exmby * exmby; out[7] = eypbx * eypbx;
    out[8] = jzb; out[9] = ez * jzb;
}

// AdaptiveTimeStep::GatherMinUzSlice (src/utils/AdaptiveTimeStep.cpp:121-141): one particle's
// {uz / c (for the minimum), w, w uz / c, w uz^2 / c^2}; false for invalid particles
HPB_HD bool adaptive_uz_terms(bool valid, double uz_c, double w, double clight_inv, double t[4])
{
    if (!valid) return false;
    const double uz = uz_c * clight_inv;
    t[0] = uz; t[1] = w; t[2] = w * uz; t[3] = w * uz * uz;
    return true;
}

// out[23]: w, w x, w x^2, w y, w y^2, w z, w z^2, w ux, w ux^2, w uy, w uy^2, w uz, w uz^2, w x ux,
// w y uy, w z uz, w x uy, w y ux, w ux/uz, w uy/uz, w gamma, w gamma^2, 1   (u = proper velocity / c).
// Returns false (nothing to add) for invalid particles and particles outside the in-situ radius.
HPB_HD bool insitu_beam_terms(bool valid, double x, double y, double z, double ux_c, double uy_c,
                              double uz_c, double w, double clight_inv, double radius_sq, double out[23])
{
    const double ux = ux_c * clight_inv, uy = uy_c * clight_inv, uz = uz_c * clight_inv;
    const double uz_inv = uz == 0. ? 0. : 1. / uz;
    if (!valid || x * x + y * y > radius_sq) return false;
    const double gamma = sqrt(1.0 + ux * ux + uy * uy + uz * uz);
    out[0] = w;
    out[1] = w * x;   out[2] = w * x * x;
    out[3] = w * y;   out[4] = w * y * y;
    out[5] = w * z;   out[6] = w * z * z;
    out[7] = w * ux;  out[8] = w * ux * ux;
    out[9] = w * uy;  out[10] = w * uy * uy;
    out[11] = w * uz; out[12] = w * uz * uz;
    out[13] = w * x * ux; out[14] = w * y * uy; out[15] = w * z * uz;
    out[16] = w * x * uy; out[17] = w * y * ux;
    out[18] = w * ux * uz_inv; out[19] = w * uy * uz_inv;
    out[20] = w * gamma; out[21] = w * gamma * gamma;
    out[22] = 1.;
    return true;
}
