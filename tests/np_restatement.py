"""TEST INFRASTRUCTURE ONLY -- a second, independent restatement of the reference's shallow-water step.

Written from the Fortran sources (not from oracle/sw_oracle.c) as whole-array NumPy expressions, one
function per reference kernel, for ONE block that covers the whole basin.  Its only purpose is to pin
the C oracle: tests/test_oracle.py runs both on the same inputs and compares every array bitwise.

Why NumPy mirrors the Fortran arithmetic exactly:
  * a binary operation on two float32 arrays is evaluated in float32 and one on float32 x float64
    promotes the float32 operand first -- Fortran's kind rules for real(4) / real(8) operands;
  * every ufunc call rounds once (no fused multiply-add, no reassociation);
  * expressions below keep the reference's left-to-right order and its parentheses;
  * double-precision literals (2.0d0, 4.0d0) are np.float64 scalars, FreeFallAcc is np.float32(9.8)
    (shared/constants.f90:11-23), so NumPy's weak Python scalars never change a kind.
Masked stores `if (mask(m,n) > 0.5) a(m,n) = ...` become np.where over the loop range; no kernel reads
an array it writes at a neighbouring cell, so whole-range evaluation equals the loop nest.

Arrays are (ny, nx) with element (m, n) of the reference's A(1:nx, 1:ny) at [n-1, m-1].
"""
import numpy as np

D2, D4 = np.float64(2.0), np.float64(4.0)
FREE_FALL_ACC = np.float32(9.8)


class Block:
    """Loop ranges of a single block: nx_start = ny_start = 3, nx_end = nx-2, ny_end = ny-2."""

    def __init__(self, nx, ny):
        self.nx, self.ny = nx, ny
        self.S = (3, nx - 2, 3, ny - 2)                    # nx_start, nx_end, ny_start, ny_end
        self.Splus = (2, nx - 1, 2, ny - 1)                # start-1 .. end+1
        self.Sminus = (2, nx - 2, 2, ny - 2)               # start-1 .. end


def view(rng):
    """Returns R(a, dm, dn): the window of `a` holding a(m+dm, n+dn) for (m, n) over the loop range."""
    m0, m1, n0, n1 = rng

    def R(a, dm=0, dn=0):
        return a[n0 - 1 + dn:n1 + dn, m0 - 1 + dm:m1 + dm]
    return R


def store(R, mask, dst, val):
    w = R(dst)
    w[...] = np.where(R(mask) > 0.5, val, w)


def f64(a):
    return a.astype(np.float64)


# kernel/shallow_water/vel_ssh.f90:69-106
def sw_update_ssh(b, tau, lu, dx, dy, dxh, dyh, hhu, hhv, sshn, sshp, ubrtr, vbrtr):
    R = view(b.S)
    div = (R(ubrtr) * R(hhu) * R(dyh) - R(ubrtr, -1) * R(hhu, -1) * R(dyh, -1)
           + R(vbrtr) * R(hhv) * R(dxh) - R(vbrtr, 0, -1) * R(hhv, 0, -1) * R(dxh, 0, -1))
    store(R, lu, sshn, R(sshp) + D2 * tau * (-div / (R(dx) * R(dy))))


def _to_u(R, q, lu, dx, dy, dxt, dyh):
    slu = f64(R(lu) + R(lu, 1))
    return (R(q) * R(dx) * R(dy) * f64(R(lu)) + R(q, 1) * R(dx, 1) * R(dy, 1) * f64(R(lu, 1))) / slu / R(dxt) / R(dyh)


def _to_v(R, q, lu, dx, dy, dxh, dyt):
    slu = f64(R(lu) + R(lu, 0, 1))
    return (R(q) * R(dx) * R(dy) * f64(R(lu)) + R(q, 0, 1) * R(dx, 0, 1) * R(dy, 0, 1) * f64(R(lu, 0, 1))) / slu / R(dxh) / R(dyt)


def _to_h(R, q, lu, dx, dy, dxb, dyb):
    slu = f64(R(lu) + R(lu, 1) + R(lu, 0, 1) + R(lu, 1, 1))
    return (R(q) * R(dx) * R(dy) * f64(R(lu)) + R(q, 1) * R(dx, 1) * R(dy, 1) * f64(R(lu, 1))
            + R(q, 0, 1) * R(dx, 0, 1) * R(dy, 0, 1) * f64(R(lu, 0, 1))
            + R(q, 1, 1) * R(dx, 1, 1) * R(dy, 1, 1) * f64(R(lu, 1, 1))) / slu / R(dxb) / R(dyb)


# kernel/shallow_water/depth.f90:14-99
def hh_init(b, ffs, lu, llu, llv, luh, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb,
            hq, hqp, hqn, hu, hup, hun, hv, hvp, hvn, hh, hhp, hhn, sh, shp, h_r):
    hq[...] = h_r + sh * np.float64(ffs)
    hqp[...] = h_r + shp * np.float64(ffs)
    hqn[...] = h_r
    R = view(b.Sminus)
    with np.errstate(divide="ignore", invalid="ignore"):
        for q, ou, ov, oh in ((hq, hu, hv, hh), (hqp, hup, hvp, hhp), (hqn, hun, hvn, hhn)):
            store(R, llu, ou, _to_u(R, q, lu, dx, dy, dxt, dyh))
            store(R, llv, ov, _to_v(R, q, lu, dx, dy, dxh, dyt))
            store(R, luh, oh, _to_h(R, q, lu, dx, dy, dxb, dyb))


# kernel/shallow_water/depth.f90:101-162
def hh_update(b, lu, llu, llv, luh, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, hqn, hun, hvn, hhn, sh, h_r):
    hqn[...] = h_r + sh
    R = view(b.Sminus)
    with np.errstate(divide="ignore", invalid="ignore"):
        store(R, llu, hun, _to_u(R, hqn, lu, dx, dy, dxt, dyh))
        store(R, llv, hvn, _to_v(R, hqn, lu, dx, dy, dxh, dyt))
        store(R, luh, hhn, _to_h(R, hqn, lu, dx, dy, dxb, dyb))


def _asselin(R, mask, ts, x, xn, xp):
    store(R, mask, xp, R(x) + ts * (R(xn) - D2 * R(x) + R(xp)) / D2)
    store(R, mask, x, R(xn))


# kernel/shallow_water/depth.f90:164-211
def hh_shift(b, ts, lu, llu, llv, luh, hq, hqp, hqn, hu, hup, hun, hv, hvp, hvn, hh, hhp, hhn):
    R = view(b.Splus)
    _asselin(R, llu, ts, hu, hun, hup)
    _asselin(R, llv, ts, hv, hvn, hvp)
    _asselin(R, lu, ts, hq, hqn, hqp)
    _asselin(R, luh, ts, hh, hhn, hhp)


# kernel/shallow_water/vel_ssh.f90:197-245
def sw_next_step(b, ts, lu, lcu, lcv, ssh, sshn, sshp, u, un, up, v, vn, vp):
    R = view(b.Splus)
    _asselin(R, lu, ts, ssh, sshn, sshp)
    _asselin(R, lcu, ts, u, un, up)
    _asselin(R, lcv, ts, v, vn, vp)


# kernel/shallow_water/vel_ssh.f90:247-281
def uv_trans_vort(b, luu, dxt, dyt, dxb, dyb, u, v, vort):
    R = view(b.S)
    val = ((R(v, 1) * R(dyt, 1) - R(v) * R(dyt))
           - (R(u, 0, 1) * R(dxt, 0, 1) - R(u) * R(dxt))
           - ((R(v, 1) - R(v)) * R(dyb) - (R(u, 0, 1) - R(u)) * R(dxb)))
    store(R, luu, vort, val)


# kernel/shallow_water/vel_ssh.f90:283-373
def uv_trans(b, lcu, lcv, luu, dxh, dyh, u, v, vort, hu, hv, hh, RHSx, RHSy):
    R = view(b.S)

    def uf(dm=0, dn=0):
        return R(u, dm, dn) * R(dyh, dm, dn) * R(hu, dm, dn)

    def vf(dm=0, dn=0):
        return R(v, dm, dn) * R(dxh, dm, dn) * R(hv, dm, dn)

    fx_p = (uf() + uf(1)) / D2 * (R(u) + R(u, 1)) / D2
    fx_m = (uf() + uf(-1)) / D2 * (R(u) + R(u, -1)) / D2
    fy_p = (vf() + vf(1)) / D2 * (R(u, 0, 1) + R(u)) / D2 * f64(R(luu))
    fy_m = (vf(0, -1) + vf(1, -1)) / D2 * (R(u, 0, -1) + R(u)) / D2 * f64(R(luu, 0, -1))
    x = (-(fx_p - fx_m + fy_p - fy_m)
         + (R(vort) * R(hh) * (R(v, 1) + R(v)) + R(vort, 0, -1) * R(hh, 0, -1) * (R(v, 1, -1) + R(v, 0, -1))) / D4)

    fy_p = (vf() + vf(0, 1)) / D2 * (R(v) + R(v, 0, 1)) / D2
    fy_m = (vf() + vf(0, -1)) / D2 * (R(v) + R(v, 0, -1)) / D2
    fx_p = (uf() + uf(0, 1)) / D2 * (R(v, 1) + R(v)) / D2
    fx_m = (uf(-1) + uf(-1, 1)) / D2 * (R(v, -1) + R(v)) / D2
    y = (-(fx_p - fx_m + fy_p - fy_m)
         - (R(vort) * R(hh) * (R(u, 0, 1) + R(u)) + R(vort, -1) * R(hh, -1) * (R(u, -1, 1) + R(u, -1))) / D4)
    store(R, lcu, RHSx, x)
    store(R, lcv, RHSy, y)


# kernel/shallow_water/mixing.f90:14-58
def stress_components(b, lu, luu, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, u, v, str_t, str_s):
    R = view(b.S)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (R(dy) / R(dx) * (R(u) / R(dyh) - R(u, -1) / R(dyh, -1))
             - R(dx) / R(dy) * (R(v) / R(dxh) - R(v, 0, -1) / R(dxh, 0, -1)))
        s = (R(dxb) / R(dyb) * (R(u, 0, 1) / R(dxt, 0, 1) - R(u) / R(dxt))
             + R(dyb) / R(dxb) * (R(v, 1) / R(dyt, 1) - R(v) / R(dyt)))
    store(R, lu, str_t, t)
    store(R, luu, str_s, s)


# kernel/shallow_water/vel_ssh.f90:375-452
def uv_diff2(b, lcu, lcv, dx, dy, dxt, dyt, dxh, dyh, dxb, dyb, mu, str_t, str_s, hq, hh, RHSx, RHSy):
    R = view(b.S)

    def sq(a, dm=0, dn=0):                                   # real(4) ** 2
        return R(a, dm, dn) * R(a, dm, dn)

    with np.errstate(divide="ignore", invalid="ignore"):
        muh_p = (R(mu) + R(mu, 1) + R(mu, 0, 1) + R(mu, 1, 1)) / D4
        muh_m = (R(mu) + R(mu, 1) + R(mu, 0, -1) + R(mu, 1, -1)) / D4
        x = ((sq(dy, 1) * R(mu, 1) * R(hq, 1) * R(str_t, 1) - sq(dy) * R(mu) * R(hq) * R(str_t)) / R(dyh)
             + (sq(dxb) * muh_p * R(hh) * R(str_s) - sq(dxb, 0, -1) * muh_m * R(hh, 0, -1) * R(str_s, 0, -1)) / R(dxt))
        muh_m = (R(mu) + R(mu, -1) + R(mu, 0, 1) + R(mu, -1, 1)) / D4
        y = (-(sq(dx, 0, 1) * R(mu, 0, 1) * R(hq, 0, 1) * R(str_t, 0, 1) - sq(dx) * R(mu) * R(hq) * R(str_t)) / R(dxh)
             + (sq(dyb) * muh_p * R(hh) * R(str_s) - sq(dyb, -1) * muh_m * R(hh, -1) * R(str_s, -1)) / R(dyt))
    store(R, lcu, RHSx, x)
    store(R, lcv, RHSy, y)


# kernel/shallow_water/vel_ssh.f90:108-195
def sw_update_uv(b, tau, lcu, lcv, dxt, dyt, dxh, dyh, dxb, dyb, hhu, hhun, hhup, hhv, hhvn, hhvp, hhh, ssh,
                 u, un, up, v, vn, vp, rdis, rlh_s, RHSx, RHSy, RHSx_adv, RHSy_adv, RHSx_dif, RHSy_dif):
    R = view(b.S)

    def cor(dm, dn, w):
        return R(rlh_s, dm, dn) * R(hhh, dm, dn) * R(dxb, dm, dn) * R(dyb, dm, dn) * w

    with np.errstate(divide="ignore", invalid="ignore"):
        bp = R(hhun) * R(dxt) * R(dyh) / D2 / tau
        bp0 = R(hhup) * R(dxt) * R(dyh) / D2 / tau
        slx = -(FREE_FALL_ACC * (R(ssh, 1) - R(ssh)) * R(dyh) * R(hhu))
        grx = (R(RHSx) + slx + R(RHSx_dif) + R(RHSx_adv)
               - (R(rdis) + R(rdis, 1)) / D2 * R(up) * R(dxt) * R(dyh) * R(hhu)
               + (cor(0, 0, R(v, 1) + R(v)) + cor(0, -1, R(v, 1, -1) + R(v, 0, -1))) / D4)
        new_u = (R(up) * bp0 + grx) / bp

        bp = R(hhvn) * R(dyt) * R(dxh) / D2 / tau
        bp0 = R(hhvp) * R(dyt) * R(dxh) / D2 / tau
        sly = -(FREE_FALL_ACC * (R(ssh, 0, 1) - R(ssh)) * R(dxh) * R(hhv))
        gry = (R(RHSy) + sly + R(RHSy_dif) + R(RHSy_adv)
               - (R(rdis) + R(rdis, 0, 1)) / D2 * R(vp) * R(dxh) * R(dyt) * R(hhv)
               - (cor(0, 0, R(u, 0, 1) + R(u)) + cor(-1, 0, R(u, -1, 1) + R(u, -1))) / D4)
        new_v = (R(vp) * bp0 + gry) / bp
    store(R, lcu, un, new_u)
    store(R, lcv, vn, new_v)


# kernel/shallow_water/vel_ssh.f90:40-67
def check_ssh_err(b, lu, ssh):
    R = view(b.S)
    sea = R(lu) > 0.5
    ok = (R(ssh) < 10000.0) & (R(ssh) > -10000.0)
    return int(np.count_nonzero(sea & ~ok))


# kernel/tracer/leapfrog_tracer.f90:13-98
def tran_diff_fluxes(b, lcu, lcv, dxt, dyt, dxh, dyh, hhu, hhv, ff, uu, vv, mu, factor_mu, flux_x, flux_y):
    R = view(b.S)
    with np.errstate(divide="ignore", invalid="ignore"):
        mu_1d = (R(mu) + R(mu, 1)) / D2 * np.float64(factor_mu) * R(dyh) / R(dxt)
        diff = mu_1d * R(hhu) * (R(ff, 1) - R(ff))
        adv = -(R(uu) * R(hhu) * R(dyh) * (R(ff) + R(ff, 1)) / D2)
        store(R, lcu, flux_x, adv + diff + np.float64(0.0))
        mu_1d = (R(mu) + R(mu, 0, 1)) / D2 * np.float64(factor_mu) * R(dxh) / R(dyt)
        diff = mu_1d * R(hhv) * (R(ff, 0, 1) - R(ff))
        adv = -(R(vv) * R(hhv) * R(dxh) * (R(ff) + R(ff, 0, 1)) / D2)
        store(R, lcv, flux_y, adv + diff + np.float64(0.0))


# kernel/tracer/leapfrog_tracer.f90:100-141
def tran_diff_tracer(b, lu, dx, dy, tau, hhqn, hhqp, flux_x, flux_y, ffp, ffn):
    R = view(b.S)
    with np.errstate(divide="ignore", invalid="ignore"):
        bp = R(hhqn) * R(dx) * R(dy) / tau / D2
        bp0 = R(hhqp) * R(dx) * R(dy) / tau / D2
        rhs = R(flux_x) - R(flux_x, -1) + R(flux_y) - R(flux_y, 0, -1)
        store(R, lu, ffn, (bp0 * R(ffp) + rhs) / bp)


# kernel/tracer/leapfrog_tracer.f90:143-170
def tracer_next_step(b, ts, lu, ffn, ffp, ff):
    R = view(b.Splus)
    _asselin(R, lu, ts, ff, ffn, ffp)


F8 = ("ssh", "sshn", "sshp", "ubrtr", "ubrtrn", "ubrtrp", "vbrtr", "vbrtrn", "vbrtrp", "RHSx", "RHSy", "RHSx_adv",
      "RHSy_adv", "RHSx_dif", "RHSy_dif", "mu", "str_t", "str_s", "vort", "hhq_rest", "hhq", "hhq_p", "hhq_n", "hhu",
      "hhu_p", "hhu_n", "hhv", "hhv_p", "hhv_n", "hhh", "hhh_p", "hhh_n")
F8_TRACER = ("flux_x", "flux_y", "ff1", "ff1n", "ff1p")
F4 = ("lu", "luu", "luh", "lcu", "lcv", "llu", "llv", "dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb", "rlh_s",
      "r_diss")


class NumpyModel:
    """ocean_data + grid_data of one block and the algorithm layer on top of the kernels above
    (control/shallow_water/shallow_water.f90:22-94, control/tracer.f90:33-62; arguments bound as in
    interface/shallow_water/sw_interface.f90:42-403 and interface/tracer/tracer_interface.f90:28-96)."""

    def __init__(self, fields, *, full_free_surface=1, trans_terms=1, ksw_lat=1, time_smooth=0.5, use_tracers=0):
        self.f = {k: np.array(v, copy=True) for k, v in fields.items()}
        ny, nx = self.f["ssh"].shape
        self.b = Block(nx, ny)
        self.ffs, self.trans, self.lat, self.tracers = full_free_surface, trans_terms, ksw_lat, use_tracers
        self.ts = np.float64(time_smooth)
        self.bad = 0

    def g(self, *names):
        return [self.f[n] for n in names]

    def step(self, tau, nsteps=1):
        tau = np.float64(tau)
        b, g, ts = self.b, self.g, self.ts
        metrics8 = ("dx", "dy", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb")
        for _ in range(nsteps):
            sw_update_ssh(b, tau, *g("lu", "dx", "dy", "dxh", "dyh", "hhu", "hhv", "sshn", "sshp", "ubrtr", "vbrtr"))
            if self.ffs > 0:
                hh_update(b, *g("lu", "llu", "llv", "luh", *metrics8, "hhq_n", "hhu_n", "hhv_n", "hhh_n", "ssh", "hhq_rest"))
            if self.trans > 0:
                uv_trans_vort(b, *g("luu", "dxt", "dyt", "dxb", "dyb", "ubrtr", "vbrtr", "vort"))
                uv_trans(b, *g("lcu", "lcv", "luu", "dxh", "dyh", "ubrtr", "vbrtr", "vort", "hhu", "hhv", "hhh",
                               "RHSx_adv", "RHSy_adv"))
            if self.lat > 0:
                stress_components(b, *g("lu", "luu", *metrics8, "ubrtrp", "vbrtrp", "str_t", "str_s"))
                uv_diff2(b, *g("lcu", "lcv", *metrics8, "mu", "str_t", "str_s", "hhq", "hhh", "RHSx_dif", "RHSy_dif"))
            sw_update_uv(b, tau, *g("lcu", "lcv", "dxt", "dyt", "dxh", "dyh", "dxb", "dyb", "hhu", "hhu_n", "hhu_p", "hhv",
                                    "hhv_n", "hhv_p", "hhh", "ssh", "ubrtr", "ubrtrn", "ubrtrp", "vbrtr", "vbrtrn",
                                    "vbrtrp", "r_diss", "rlh_s", "RHSx", "RHSy", "RHSx_adv", "RHSy_adv", "RHSx_dif",
                                    "RHSy_dif"))
            sw_next_step(b, ts, *g("lu", "lcu", "lcv", "ssh", "sshn", "sshp", "ubrtr", "ubrtrn", "ubrtrp", "vbrtr",
                                   "vbrtrn", "vbrtrp"))
            depth12 = ("hhq", "hhq_p", "hhq_n", "hhu", "hhu_p", "hhu_n", "hhv", "hhv_p", "hhv_n", "hhh", "hhh_p", "hhh_n")
            if self.ffs > 0:
                hh_shift(b, ts, *g("lu", "llu", "llv", "luh", *depth12))
                hh_init(b, self.ffs, *g("lu", "llu", "llv", "luh", *metrics8, *depth12, "ssh", "sshp", "hhq_rest"))
            self.bad += check_ssh_err(b, *g("lu", "ssh"))
            if self.tracers > 0:
                tran_diff_fluxes(b, *g("lcu", "lcv", "dxt", "dyt", "dxh", "dyh", "hhu", "hhv", "ff1", "ubrtr", "vbrtr", "mu"),
                                 1.0, *g("flux_x", "flux_y"))
                tran_diff_tracer(b, *g("lu", "dx", "dy"), tau, *g("hhq_n", "hhq_p", "flux_x", "flux_y", "ff1p", "ff1n"))
                tracer_next_step(b, ts, *g("lu", "ff1n", "ff1p", "ff1"))
