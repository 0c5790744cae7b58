// rydcz_prologue.cuh -- the scalar prologue of simulate_CZ_gate, one thread per parameter point.
//
// Restates RG/simulation.py:2796-3355 (and what it calls in laser_physics, trap_physics,
// protocols, noise_models part 1) for the three protocols whose schedules are closed-form:
// square LP, JP bang-bang, smooth sinusoidal JP.  Species- and n-level constants (C6(n), dipoles,
// lifetimes, polarizabilities ...) are computed once per call on the host by the Python mirror
// and arrive in the `model` array; everything that depends on the per-point apparatus values
// (laser powers, spacing, temperature, tweezer, B field, protocol parameters) is computed here and
// written straight into the evaluator's point / segment rows -- the "assembly from pulse
// parameters" stage of the path, so a sweep or an optimiser generation goes host -> device as ~40
// doubles per point and never materialises Hamiltonians or rate tables on the host.
//
// Plain C++ when compiled by g++ (tests/hostsim), like rydcz_core.cuh.
#pragma once
#include <math.h>

#include "rydcz_core.cuh"

namespace rydcz {

// ---- apparatus slots (doubles); NaN = "None" (use the reference's default) --------------------
enum AppSlot {
  APP_P1 = 0, APP_W1 = 1, APP_P2 = 2, APP_W2 = 3, APP_DELTA_E = 4,
  APP_LW1 = 5, APP_LW2 = 6,          // laser linewidths (Hz), NaN = None          RG/simulation.py:2853-2860
  APP_PURITY = 7,                    // min(polarization_purity_1, _2)             RG/simulation.py:3307
  APP_TW_POWER = 8, APP_TW_WAIST = 9,
  APP_TW_WL_NM = 10,                 // tweezer_wavelength_nm, NaN = species default
  APP_TEMPERATURE = 11, APP_B_FIELD = 12, APP_NA = 13, APP_SPACING = 14,
  APP_BG_LOSS = 15,                  // background_loss_rate_hz, NaN = 1e3 default
  APP_INTENSITY_FRAC = 16,
  APP_PROTO = 20,                    // protocol block: LP  [Δ/Ω, Ωτ]
                                     //                 SJP [Ωτ, A, ω_mod/Ω, φ_off, δ/Ω]
                                     //                 JP  [Ωτ, S-1 switching times, S phases]
  APP_STRIDE = 40
};

// ---- per-call model constants (doubles; flags stored as 0/1) -----------------------------------
enum ModelSlot {
  MD_DIM = 0, MD_PROTOCOL = 1 /* 0 lp, 1 jp, 2 sjp */, MD_NSEG = 2, MD_INCLUDE_NOISE = 3,
  MD_MOTIONAL = 4, MD_DOPPLER = 5, MD_INTENSITY = 6, MD_TRAP_ON = 7, MD_CLOCK = 8,
  MD_LEAK_SHAPE = 9 /* 0 square, 1 gaussian, 2 cosine, 3 blackman, 4 drag, 5 other */,
  MD_MASS = 10, MD_ALPHA_GROUND = 11, MD_TRAP_WL = 12, MD_C6 = 13, MD_DIPOLE_1E = 14, MD_DIPOLE_ER = 15,
  MD_GAMMA_E = 16, MD_GAMMA_R = 17, MD_ALPHA_RYD_STATIC = 18, MD_OMEGA_D1 = 19, MD_K_QUAD_DEPH = 20,
  MD_ZEEMAN_K = 21 /* clock: Hz/G^2 */, MD_ZEEMAN_LIN = 22 /* non-clock: rad/s per T */,
  MD_STARK_HZ_PER_MK = 23, MD_STARK_ALT = 24 /* Hz per (W/m^2) */, MD_K_EFF = 25, MD_DELTA_LEAK = 26,
  MD_RYD_ZEEMAN_PER_T = 27, MD_BG_DEFAULT = 28,
  MD_E_BRANCH_1 = 29 /* dim 5: fraction of |e> decays that return to |1> */,
  MD_LS_COMP = 30 /* dim 5: 1 = light shifts compensated */,
  MD_AREA_CORRECTION = 31 /* shaped LP: square area / envelope area (RG/pulse_shaping.py:795-842), 1 for square */,
  MD_STRIDE = 32
};

// scalars reported back per point (SimulationResult fields that are not kernel inputs)
enum ScalarSlot { SC_OMEGA = 0, SC_V = 1, SC_TAU_TOTAL = 2, SC_TAU_SINGLE = 3, SC_DELTA = 4, SC_XI_RE = 5,
                  SC_XI_IM = 6, SC_OMEGA_TAU = 7, SC_STRIDE = 8 };

namespace phys {
constexpr double HBAR = 1.054571817e-34, EPS0 = 8.8541878128e-12, C = 299792458.0, KB = 1.380649e-23;
constexpr double PI = 3.141592653589793238462643383279;
}  // namespace phys

// LP (Δ/Ω, Ωτ) against V/Ω: log-linear interpolation, clamped to [10, 1000]  (RG/protocols.py:369-379, 614-644)
RYDCZ_HD void lp_table(double v, double& d, double& o) {
  const double key[7] = {10, 25, 50, 100, 200, 500, 1000};
  const double dd[7] = {0.340, 0.360, 0.370, 0.375, 0.377, 0.3773, 0.37737};
  const double oo[7] = {4.45, 4.35, 4.32, 4.30, 4.293, 4.2927, 4.29268};
  v = fmin(fmax(v, 10.0), 1000.0);
  int hi = 0;
  while (hi < 6 && key[hi] < v) ++hi;
  const int lo = (key[hi] == v || hi == 0) ? hi : hi - 1;
  const double t = (lo == hi) ? 0.0 : (log(v) - log(key[lo])) / (log(key[hi]) - log(key[lo]));
  d = dd[lo] + t * (dd[hi] - dd[lo]);
  o = oo[lo] + t * (oo[hi] - oo[lo]);
}

// RG/trap_physics.py:1169-1203
RYDCZ_HD double thermal_dephasing(double dv, double V0, double Omega) {
  if (Omega <= 0.0) Omega = 2 * phys::PI * 5e6;
  const double aO = fabs(Omega), aV = fabs(V0), r = aV / aO, dv2 = dv * dv;
  const double g_weak = (dv2 * (r * r)) * aO / (2 * phys::PI);
  const double sup = (aO / aV) * (aO / aV);
  const double g_strong = (dv2 * sup) * aO / (2 * phys::PI);
  double g;
  if (r < 3) g = g_weak;
  else if (r > 10) g = g_strong;
  else {
    double x = fmin(fmax((r - 3) / 7, 0.0), 1.0);
    const double s = 3 * (x * x) - 2 * (x * x * x);
    g = g_weak * (1 - s) + g_strong * s;
  }
  return fmin(g, 10e6);
}

// RG/trap_physics.py:944-995, 1040-1057
RYDCZ_HD double antitrap_loss_rate(double t_gate, double U0, double alpha_ratio, double mass, double waist,
                                   double T) {
  const double t_ryd = 0.3 * t_gate;
  const double w_trap = sqrt(4 * U0 / (mass * (waist * waist)));
  const double v_th = sqrt(phys::KB * T / mass);
  const double capture = 2.0 * waist;
  const double w_anti = sqrt(4 * alpha_ratio * U0 / (mass * (waist * waist)));
  double P = 0.0;
  if (w_anti > 0 && t_ryd > 0) {
    const double s0 = sqrt(phys::KB * T / (mass * (w_trap * w_trap)));
    const double sp = s0 * cosh(w_anti * t_ryd), sv = (v_th / w_anti) * sinh(w_anti * t_ryd);
    const double sig = sqrt(sp * sp + sv * sv);
    if (sig > 0) { const double q = capture / sig; P = 1.0 - exp(-(q * q) / 2); }
  }
  P = fmin(fmax(P, 0.0), 1.0);
  double g;
  if (P >= 0.99) g = 5.0 / t_gate;
  else if (P > 0) g = -log(1 - P) / t_gate;
  else g = 0.0;
  const double cap = t_gate > 0 ? 1.0 / t_gate : 1e6;
  return fmin(g, cap);
}

// RG/noise_models.py:807-849
RYDCZ_HD double leakage_rate(double Omega, double Dleak, int shape, double tau, double gamma_r) {
  if (fabs(Dleak) < 1e-6 || fabs(Omega) < 1e-6) return 0.0;
  const double x = Dleak * tau / (2 * phys::PI);
  double S;
  if (fabs(x) < 1e-10) S = 1.0;
  else if (shape == 0) { const double q = sin(phys::PI * x) / (phys::PI * x); S = q * q; }
  else if (shape == 1) { const double q = Dleak * tau / 8; S = exp(-(q * q)); }
  else if (shape == 2) {
    if (fabs(fabs(x) - 0.5) < 1e-10) S = 0.25;
    else { const double q = sin(phys::PI * x) / (phys::PI * x * (1 - x * x)); S = q * q; }
  } else if (shape == 3) S = exp(-3 * fabs(x)) * 0.1;
  else if (shape == 4) { const double q = Dleak * tau / 8; S = exp(-(q * q)) * 0.1; }
  else { const double q = sin(phys::PI * x) / (phys::PI * x + 1e-10); S = q * q; }
  S = fmin(fmax(S, 0.0), 1.0);
  const double f = Omega / Dleak;
  return (f * f) * gamma_r * S;
}

// One point: apparatus row -> evaluator point row (PT_STRIDE doubles; PT5_STRIDE for dim 5), segment
// rows, reported scalars.  `seg` must have room for MD_NSEG rows.  Returns the number of segments.
// The apparatus block is slot-major: slot k of point i is app_base[k * n_points + i] (contiguous
// columns: cheap to fill on the host, coalesced on the device); `app` below is &app_base[i].
struct AppView {
  const double* base;
  long long stride;
  RYDCZ_HD double operator[](int k) const { return base[(long long)k * stride]; }
};

// Long schedules (smooth JP: 300 rows, shaped LP: 998) are written by one thread per ROW on the device: the point's
// thread leaves the handful of numbers a row needs in `wave` and schedule_row below turns them into row s.  (One
// thread per point writing 300 rows one after the other -- 300 dependent sincos, 9.6 KB of scattered stores -- was
// 0.22 ms of a 0.96 ms generation of 8192 smooth-JP candidates.)  The serial path calls the same function.
constexpr int WAVE_STRIDE = 10;
enum WaveSlot { WV_KIND = 0 /* 0 none, 1 shaped LP, 2 smooth JP */, WV_N = 1, WV_T = 2, WV_AMP = 3, WV_DET = 4,
                WV_A = 5, WV_B = 6, WV_C = 7, WV_ROWS = 8 };
RYDCZ_HD void schedule_row(const double* __restrict__ wave, long long wstride, int s, double* __restrict__ row) {
  const double PI_ = 3.14159265358979323846;
  const int kind = (int)wave[WV_KIND * wstride];
  const double T = wave[WV_T * wstride], amp = wave[WV_AMP * wstride], det = wave[WV_DET * wstride];
  const double a_ = wave[WV_A * wstride], b_ = wave[WV_B * wstride], c_ = wave[WV_C * wstride];
  if (kind == 2) {
    // smooth JP: T = dts, amp = Omega, a_ = A, b_ = omega_mod, c_ = phi
    const double t_mid = s * T + T / 2;
    const double ph = a_ * cos(b_ * t_mid - c_);
    row[0] = T; row[1] = amp * cos(ph); row[2] = amp * sin(ph); row[3] = det;
  } else {
    // shaped LP: T = tau_single, amp = peak, a_ + i b_ = xi, c_ = shape; rows 0 .. n-2 first pulse, then the second
    const int n = (int)wave[WV_N * wstride];
    const bool second = s >= n - 1;
    const int k = second ? s - (n - 1) : s;
    const double step = T / (double)(n - 1), dt = T / (double)n;
    const double t0 = (double)k * step, t1 = (k + 1 == n - 1) ? T : (double)(k + 1) * step;   // np.linspace
    const double t_mid = (t0 + t1) / 2;
    double env = 1.0;
    if ((int)c_ == 2) { const double sn = sin(PI_ * t_mid / T); env = sn * sn; }
    const double a = amp * env;
    row[0] = dt; row[1] = second ? a * a_ : a; row[2] = second ? a * b_ : 0.0; row[3] = det;
  }
}

// `wave` (may be null): WAVE_STRIDE doubles of this point, element k at wave[k * wstride].  Given, a long schedule is
// NOT written here: the caller runs schedule_row for rows 0 .. wave[WV_ROWS] - 1 (wave[WV_KIND] = 0: written here).
RYDCZ_HD int prologue_point(const double* __restrict__ md, const AppView app,
                            double* __restrict__ pt, double* __restrict__ seg, double* __restrict__ sc,
                            double* __restrict__ wave = nullptr, long long wstride = 1) {
  using namespace phys;
  if (wave) { wave[WV_KIND * wstride] = 0.0; wave[WV_ROWS * wstride] = 0.0; }
  const int protocol = (int)md[MD_PROTOCOL], dim = (int)md[MD_DIM], S = (int)md[MD_NSEG];
  const double P1 = app[APP_P1], w1 = app[APP_W1], P2 = app[APP_P2], w2 = app[APP_W2], De = app[APP_DELTA_E];
  const double lw1 = app[APP_LW1], lw2 = app[APP_LW2];
  double lw;
  if (lw1 == lw1 && lw2 == lw2) lw = sqrt(lw1 * lw1 + lw2 * lw2);
  else if (lw1 == lw1) lw = lw1;
  else if (lw2 == lw2) lw = lw2;
  else lw = 1000.0;
  const double T = app[APP_TEMPERATURE], Bf = app[APP_B_FIELD];
  const double trap_wl = (app[APP_TW_WL_NM] == app[APP_TW_WL_NM]) ? app[APP_TW_WL_NM] * 1e-9 : md[MD_TRAP_WL];
  const double R = app[APP_SPACING] * (trap_wl / (2 * app[APP_NA]));
  // Rabi chain
  const double E1 = sqrt(2 * (2 * P1 / (PI * (w1 * w1))) / (EPS0 * C));
  const double E2 = sqrt(2 * (2 * P2 / (PI * (w2 * w2))) / (EPS0 * C));
  const double Om1 = md[MD_DIPOLE_1E] * E1 / HBAR, Om2 = md[MD_DIPOLE_ER] * E2 / HBAR;
  const double Om = Om1 * Om2 / (2 * De);
  const double V = md[MD_C6] / pow(R, 6.0);
  const double VoO = Om > 0 ? V / Om : INFINITY;
  // protocol parameters
  const AppView pp{app.base + (long long)APP_PROTO * app.stride, app.stride};
  double o_t, d_o, tau_single, tau_total, Delta_gate;
  if (protocol == 0) {
    double td, to;
    lp_table(VoO, td, to);
    d_o = (pp[0] == pp[0]) ? pp[0] : td;
    o_t = (pp[1] == pp[1]) ? pp[1] : to;
    tau_single = o_t / Om; tau_total = 2 * tau_single; Delta_gate = d_o * Om;
  } else if (protocol == 1) {
    o_t = (pp[0] == pp[0]) ? pp[0] : 22.08;
    tau_single = o_t / Om; tau_total = tau_single; Delta_gate = 0.0; d_o = 0.0;
  } else {
    o_t = (pp[0] == pp[0]) ? pp[0] : 10.09;
    d_o = (pp[4] == pp[4]) ? pp[4] : 0.0205;
    tau_single = o_t / Om; tau_total = tau_single; Delta_gate = 0.0;
  }
  // trap-dependent noise (RG/trap_physics.py:1704-1848)
  const double mass = md[MD_MASS], twp = app[APP_TW_POWER], tww = app[APP_TW_WAIST];
  const double U0 = fabs(md[MD_ALPHA_GROUND]) * (2 * twp / (PI * (tww * tww))) / (2 * EPS0 * C);
  const double w_r = sqrt(4 * U0 / (mass * (tww * tww)));
  const double sigma_r = sqrt(KB * T / (mass * (w_r * w_r)));
  const double dV = 6 * (sqrt(2.0) * sigma_r) / R;
  const double g_thermal = thermal_dephasing(dV, V, Om);
  const double Ge = md[MD_GAMMA_E];
  const double g_scatter = (Om1 > 0 && De > 0) ? Ge * ((Om1 / 2) * (Om1 / 2)) / (De * De + (Ge / 2) * (Ge / 2)) : 0.0;
  const double wl_nm = trap_wl * 1e9;
  const double wl_m = wl_nm * 1e-9;
  const double om_trap = 2 * PI * C / wl_m, oD1 = md[MD_OMEGA_D1];
  const double a0 = md[MD_ALPHA_GROUND];
  const double a_g = (om_trap < oD1) ? a0 * (1.0 + 0.3 * (om_trap * om_trap) / (oD1 * oD1 - om_trap * om_trap)) : a0;
  const double wr = wl_m / 1064e-9;
  const double a_r = md[MD_ALPHA_RYD_STATIC] * (wr * wr);
  const double alpha_ratio = fabs(a_g) > 1e-50 ? fabs(a_r / a_g) : 0.0;
  const double g_anti_raw = (alpha_ratio > 0 && tau_total > 0)
                                ? antitrap_loss_rate(tau_total, U0, alpha_ratio, mass, tww, T) : 0.0;
  const double ratio_signed = fabs(a_g) > 1e-50 ? a_r / a_g : 0.0;
  const double enhancement = 1.0 / (1.0 + fabs(1.0 - ratio_signed));
  const double v_rms = sqrt(KB * T / mass);
  const double k_eff = md[MD_K_EFF];
  const double g_doppler = (md[MD_DOPPLER] != 0 && tau_total > 0) ? ((k_eff * v_rms) * (k_eff * v_rms)) * tau_total : 0.0;
  const double frac = app[APP_INTENSITY_FRAC];
  const double g_intensity = (md[MD_INTENSITY] != 0 && frac > 0) ? (U0 / HBAR) * frac * fmin(enhancement, 0.1) : 0.0;
  const double gamma_r = md[MD_GAMMA_R];
  // coherent shifts (RG/trap_physics.py:1908-1965, 2112-2142)
  const double Bg = Bf * 1e4;
  const double d_zee = md[MD_CLOCK] != 0 ? md[MD_ZEEMAN_K] * (Bg * Bg) * 2 * PI : md[MD_ZEEMAN_LIN] * Bf;
  double d_stark = 0.0;
  if (md[MD_TRAP_ON] != 0) {
    const double depth_mK = (U0 / KB * 1e6) / 1000;
    const double hz = depth_mK > 0 ? md[MD_STARK_HZ_PER_MK] * depth_mK
                                   : md[MD_STARK_ALT] * (2 * twp / (PI * (tww * tww)));
    d_stark = hz * 2 * PI;
  }
  // LP phase jump (RG/protocols.py:798-819)
  double xi_re = 1.0, xi_im = 0.0;
  if (protocol == 0) {
    const double aO = fabs(Om);
    if (aO >= 1e-10) {
      const double y = Delta_gate / aO, a = sqrt(y * y + 1), b = aO * tau_single * a / 2;
      const double nr = a * cos(b), ni = y * sin(b), dr = -a * cos(b), di = y * sin(b);
      const double den = dr * dr + di * di;
      if (sqrt(den) >= 1e-12) { xi_re = (nr * dr + ni * di) / den; xi_im = (ni * dr - nr * di) / den; }
    }
  }
  // rates (RG/simulation.py:3230-3334)
  double r5[7] = {0, 0, 0, 0, 0, 0, 0};   // [5], [6]: dim 5 split of the |r> loss channels (-> |0>, -> |l>)
  if (md[MD_INCLUDE_NOISE] != 0) {
    const double g_phi_laser = PI * lw;
    const double g_bg = (app[APP_BG_LOSS] == app[APP_BG_LOSS]) ? app[APP_BG_LOSS] : md[MD_BG_DEFAULT];
    const double g_motional = md[MD_MOTIONAL] != 0 ? g_thermal : 0.0;
    const double B_rms = fmax(0.01 * Bf * 1e4, 1e-3);
    const double g_zee = md[MD_CLOCK] != 0 ? 2 * PI * (2 * md[MD_K_QUAD_DEPH] * 1.0 * B_rms) : 2 * PI * (700e3 * B_rms);
    const double tf = fmin(1.0, (tau_total / 1e-6) * (tau_total / 1e-6));
    const double g_anti = g_anti_raw * 0.3 * tf;
    const double g_leak = leakage_rate(Om, md[MD_DELTA_LEAK], (int)md[MD_LEAK_SHAPE], tau_single, gamma_r);
    const double g_th_total = g_motional + g_doppler + g_intensity;
    double g_mj = 0.0;
    if (dim == 4) {
      const double eps = 1.0 - app[APP_PURITY], dz = md[MD_RYD_ZEEMAN_PER_T] * Bf;
      g_mj = fabs(dz) < 1e-10 ? (eps * eps) * fabs(Om) : (eps * eps) * (Om * Om) / fabs(dz);
    }
    r5[0] = gamma_r * 0.5;
    r5[1] = gamma_r * (1 - 0.5) + 0.0 + g_anti + g_bg + g_leak;
    r5[2] = g_phi_laser + g_th_total + g_zee;
    r5[3] = g_scatter;
    r5[4] = g_mj;
    r5[5] = gamma_r * (1 - 0.5) + 0.0;
    r5[6] = g_anti + g_bg + g_leak;
  }
  pt[PT_V] = V; pt[PT_DELTA1] = d_zee + d_stark; pt[PT_OFF_A] = 0.0; pt[PT_OFF_B] = 0.0;
  pt[PT_G_R1] = r5[0]; pt[PT_G_R0] = r5[1]; pt[PT_G_PHI] = r5[2]; pt[PT_G_SC] = r5[3]; pt[PT_G_MJ] = r5[4];
  pt[PT_SIG_DET] = k_eff * v_rms;
  pt[PT_SIG_RELR] = sqrt(2.0) * sigma_r / R;
  pt[PT_RESERVED] = 0.0;
  if (dim == 5) {
    // the intermediate state is explicit: scattering is the decay of |e>, loss
    // channels that remove the atom end in |l>
    const bool noisy5 = md[MD_INCLUDE_NOISE] != 0;
    pt[PT_G_SC] = 0.0;
    pt[PT5_OM1] = Om1; pt[PT5_DET_E] = De;
    pt[PT5_G_E1] = noisy5 ? Ge * md[MD_E_BRANCH_1] : 0.0;
    pt[PT5_G_E0] = noisy5 ? Ge * (1.0 - md[MD_E_BRANCH_1]) : 0.0;
    pt[PT_G_R0] = r5[5];
    pt[PT5_G_RL] = r5[6];
    pt[PT5_LS_COMP] = md[MD_LS_COMP];
    pt[18] = 0.0; pt[19] = 0.0;
  }
  sc[SC_OMEGA] = Om; sc[SC_V] = V; sc[SC_TAU_TOTAL] = tau_total; sc[SC_TAU_SINGLE] = tau_single;
  sc[SC_DELTA] = Delta_gate; sc[SC_XI_RE] = xi_re; sc[SC_XI_IM] = xi_im; sc[SC_OMEGA_TAU] = o_t;
  // pulse schedule
  if (protocol == 0 && S == 2) {
    seg[0] = tau_single; seg[1] = Om; seg[2] = 0.0; seg[3] = Delta_gate;
    seg[4] = tau_single; seg[5] = Om * xi_re; seg[6] = Om * xi_im; seg[7] = Delta_gate;
    return 2;
  }
  if (protocol == 0) {
    // Shaped LP (RG/simulation.py:2099-2231): two pulses of n - 1 midpoint-sampled steps on a grid of n = S / 2 + 1
    // points, each step lasting tau / n (the reference's dt: 499 steps of tau / 500), amplitude Omega x (area
    // correction) x envelope(t_mid).  The reference evaluates the envelope on a SCALAR time, so the shapes that
    // normalise by their own maximum (gaussian, blackman; RG/pulse_shaping.py:184-186, 291-293) return 1: they run
    // as square pulses of the corrected amplitude.  Only the cosine envelope sin^2(pi t / tau) really shapes.
    const int n = S / 2 + 1, shape = (int)md[MD_LEAK_SHAPE];
    const double peak = Om * md[MD_AREA_CORRECTION];
    double wv[WAVE_STRIDE];
    double* w = wave ? wave : wv;
    const long long ws = wave ? wstride : 1;
    w[WV_KIND * ws] = 1.0; w[WV_N * ws] = (double)n; w[WV_T * ws] = tau_single; w[WV_AMP * ws] = peak;
    w[WV_DET * ws] = Delta_gate; w[WV_A * ws] = xi_re; w[WV_B * ws] = xi_im; w[WV_C * ws] = (double)shape;
    w[WV_ROWS * ws] = (double)(2 * (n - 1));
    if (!wave)
      for (int k = 0; k < 2 * (n - 1); ++k) schedule_row(w, ws, k, seg + 4 * k);
    return 2 * (n - 1);
  }
  if (protocol == 1) {
    // boundaries [0] + switching + [Ωτ], over Ω; dt < 1e-18 skipped (RG/simulation.py:1859-1903)
    double prev = 0.0 / Om;
    for (int s = 0; s < S; ++s) {
      const double nxt = (s < S - 1 ? pp[1 + s] : o_t) / Om;
      double dt = nxt - prev;
      if (dt < 1e-18) dt = 0.0;
      const double ph = pp[S + s];
      seg[4 * s + 0] = dt; seg[4 * s + 1] = Om * cos(ph); seg[4 * s + 2] = Om * sin(ph); seg[4 * s + 3] = 0.0;
      prev = nxt;
    }
    return S;
  }
  // smooth sinusoidal JP (RG/simulation.py:1693-1733, 3462-3483); `x or default` quirk for A, ratio, phi
  const double A = (pp[1] == pp[1] && pp[1] != 0.0) ? pp[1] : 0.311 * PI;
  const double ratio = (pp[2] == pp[2] && pp[2] != 0.0) ? pp[2] : 1.242;
  const double phi = (pp[3] == pp[3] && pp[3] != 0.0) ? pp[3] : 4.696;
  const double doo = De > 0 ? -fabs(d_o) : fabs(d_o);
  const double dts = tau_total / S, om_mod = ratio * Om, Dsm = doo * Om;
  double wv[WAVE_STRIDE];
  double* w = wave ? wave : wv;
  const long long ws = wave ? wstride : 1;
  w[WV_KIND * ws] = 2.0; w[WV_N * ws] = (double)S; w[WV_T * ws] = dts; w[WV_AMP * ws] = Om; w[WV_DET * ws] = Dsm;
  w[WV_A * ws] = A; w[WV_B * ws] = om_mod; w[WV_C * ws] = phi; w[WV_ROWS * ws] = (double)S;
  if (!wave)
    for (int k = 0; k < S; ++k) schedule_row(w, ws, k, seg + 4 * k);
  return S;
}

}  // namespace rydcz
