"""tests/independent_modal.py -- an INDEPENDENT statement of the surface-wave eigenproblem, used only to
check the oracle's roots against physics (TEST INFRASTRUCTURE; nothing under mctomo_b200/ may import it).

It shares no formula with surfdisp96: no Dunkin compound matrix, no Haskell layer matrices, no
half-space closed forms.  The elastic equations of motion for a plane surface wave in a stack of
homogeneous layers are written as first-order systems  dy/dz = A y  (z down):

  P-SV (Rayleigh), y = (r1, r2, r3, r4) = (horizontal disp., vertical disp., tau_zx, tau_zz):
      r1' =  k r2 + r3 / mu
      r2' = -k lam/(lam+2mu) r1 + r4 / (lam+2mu)
      r3' = (k^2 zeta - w^2 rho) r1 + k lam/(lam+2mu) r4,      zeta = 4 mu (lam+mu)/(lam+2mu)
      r4' = -w^2 rho r2 - k r3
  SH (Love), y = (l1, l2) = (disp., tau_zy):   l1' = l2 / mu,   l2' = (k^2 mu - w^2 rho) l1
  fluid,     y = (r2, r4):                      r2' = (1/lam - k^2/(rho w^2)) r4,   r4' = -rho w^2 r2

and are solved numerically in multiprecision arithmetic (mpmath): the layer propagators are matrix
exponentials expm(-A h), the half-space radiation condition is "the eigenvectors of A whose eigenvalues
have negative real part", and the secular function is the determinant of the surface tractions of the two
(P-SV) or the traction of the one (SH) decaying solution carried up to the free surface.  With 50 digits
the growing exponentials that make the double-precision Thomson-Haskell form unstable are harmless."""
import mpmath as mp

mp.mp.dps = 50


def _solid_A(k, w, a, b, rho):
    mu = rho * b * b
    lam = rho * a * a - 2 * mu
    l2m = lam + 2 * mu
    zeta = 4 * mu * (lam + mu) / l2m
    return mp.matrix([[0, k, 1 / mu, 0],
                      [-k * lam / l2m, 0, 0, 1 / l2m],
                      [k * k * zeta - w * w * rho, 0, 0, k * lam / l2m],
                      [0, -w * w * rho, -k, 0]])


def _decaying(A, n):
    """Columns spanning the solutions of y' = A y that decay as z -> +inf."""
    E, ER = mp.eig(A)
    idx = sorted((i for i in range(len(E)) if mp.re(E[i]) < 0), key=lambda i: mp.re(E[i]))
    assert len(idx) == n, "phase velocity must be below the half-space body-wave velocities"
    cols = []
    for i in idx:  # fixed order (by decay rate) and fixed sign/scale (first component = +1: it never vanishes
        v = mp.matrix([mp.re(x) for x in ER[:, i]])  # below the shear velocity), so that the secular function
        cols.append(v / v[0])                        # is continuous in c
    return cols


def rayleigh_secular(c, period, thick, vp, vs, rho):
    """Secular function of the P-SV problem at phase velocity c (km/s) and period (s); the top layer may be water."""
    w = 2 * mp.pi / mp.mpf(period)
    k = w / mp.mpf(c)
    n = len(thick)
    first_solid = 1 if vs[0] == 0 else 0
    y1, y2 = _decaying(_solid_A(k, w, mp.mpf(vp[n - 1]), mp.mpf(vs[n - 1]), mp.mpf(rho[n - 1])), 2)
    for m in range(n - 2, first_solid - 1, -1):
        P = mp.expm(-_solid_A(k, w, mp.mpf(vp[m]), mp.mpf(vs[m]), mp.mpf(rho[m])) * mp.mpf(thick[m]))
        y1, y2 = P * y1, P * y2
        s = max(abs(v) for v in list(y1) + list(y2))     # positive rescaling: keeps the sign of the determinant
        y1, y2 = y1 / s, y2 / s
    if not first_solid:
        return y1[2] * y2[3] - y1[3] * y2[2]             # tau_zx = tau_zz = 0 at the free surface
    # water on top: the combination with tau_zx = 0 at the sea floor, carried through the fluid; pressure-free surface
    yc = y2[2] * y1 - y1[2] * y2
    lam = mp.mpf(rho[0]) * mp.mpf(vp[0]) ** 2
    Af = mp.matrix([[0, 1 / lam - k * k / (mp.mpf(rho[0]) * w * w)], [-mp.mpf(rho[0]) * w * w, 0]])
    top = mp.expm(-Af * mp.mpf(thick[0])) * mp.matrix([yc[1], yc[3]])
    return top[1]


def love_secular(c, period, thick, vs, rho):
    """Secular function of the SH problem; a water layer on top is ignored (it carries no shear)."""
    w = 2 * mp.pi / mp.mpf(period)
    k = w / mp.mpf(c)
    n = len(thick)
    first_solid = 1 if vs[0] == 0 else 0

    def A(m):
        mu = mp.mpf(rho[m]) * mp.mpf(vs[m]) ** 2
        return mp.matrix([[0, 1 / mu], [k * k * mu - w * w * mp.mpf(rho[m]), 0]])
    (y,) = _decaying(A(n - 1), 1)
    for m in range(n - 2, first_solid - 1, -1):
        y = mp.expm(-A(m) * mp.mpf(thick[m])) * y
        y = y / max(abs(y[0]), abs(y[1]))
    return y[1]


def count_sign_changes(f, lo, hi, n):
    """Number of sign changes of f on n equal steps of (lo, hi) -- the number of modes when n is fine enough."""
    cnt = 0
    prev = None
    for i in range(n + 1):
        v = f(lo + (hi - lo) * i / n)
        s = v > 0
        if prev is not None and s != prev:
            cnt += 1
        prev = s
    return cnt
