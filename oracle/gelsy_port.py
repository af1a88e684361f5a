"""TEST INFRASTRUCTURE (oracle) -- never imported by the shipped path.

Plain numpy restatement of LAPACK's ZGELSY, the driver behind the reference's
`scipy.linalg.lstsq(x4_q, y_q.T, lapack_driver="gelsy")` (/root/reference/fftisdf.py:108).

The algorithm lives in a third-party dependency that is not in /root/reference: LAPACK 3.x as shipped in
OpenBLAS 0.3.30 inside scipy 1.18 (the build image's versions; the reference pins none).  Its published
algorithm is restated here routine by routine:

    zgeqp3/zlaqp2  Householder QR with column pivoting on partial column norms (norm downdating with the
                   sqrt(eps) recomputation safeguard)                                   -> qrcp()
    zlarfg         complex elementary reflector                                          -> larfg()
    zlaic1         one step of incremental condition estimation (jobs 1 and 2)           -> laic1()
    zgelsy         rank = number of leading columns of R with smax*rcond <= smin, rcond = eps (scipy's
                   default `cond`), complete orthogonal factorisation [R11 R12] = [T11 0] Z (ztzrzf),
                   x = P Z^H [T11^-1 (Q^H b)(1:rank); 0]                                  -> gelsy(), gelsy_factor()

Pinned against the real LAPACK through scipy (tests/test_oracle_cpu.py): same pivots as scipy's zgeqp3 on
tie-free matrices, same rank as scipy's lstsq(gelsy) on graded / rank-deficient matrices, same solution.
`gelsy_factor` returns the factored operator the CUDA path applies (Q1, T11, E = P Z1^H), so that the device
kernels can be checked piece by piece.
"""
import numpy as np

EPS = float(np.finfo(np.float64).eps) * 0.5      # dlamch('Epsilon') = 2^-53
RCOND = float(np.finfo(np.float64).eps)          # scipy.linalg.lstsq default `cond` for gelsy
TOL3Z = float(np.sqrt(EPS))                      # zlaqp2 / zlaqps


def larfg(alpha, x):
    """zlarfg: H = I - tau v v^H with v = [1; x_out], H^H [alpha; x] = [beta; 0], beta real.
    Returns (beta, tau, x_out)."""
    xnorm = float(np.linalg.norm(x))
    alphr, alphi = float(alpha.real), float(alpha.imag)
    if xnorm == 0.0 and alphi == 0.0:
        return alpha, 0.0 + 0.0j, x
    beta = -np.copysign(np.sqrt(alphr * alphr + alphi * alphi + xnorm * xnorm), alphr)
    tau = complex((beta - alphr) / beta, -alphi / beta)
    scale = 1.0 / (alpha - beta)
    return complex(beta, 0.0), tau, x * scale


def qrcp(a):
    """zgeqp3 by the unblocked recurrence (zlaqp2).  a [m, n] complex -> (qr, tau, jpvt) in LAPACK's packed
    form (R in the upper triangle, reflector vectors below the diagonal), jpvt 0-based."""
    a = np.array(a, dtype=np.complex128, order="F")
    m, n = a.shape
    mn = min(m, n)
    jpvt = np.arange(n)
    tau = np.zeros(mn, dtype=np.complex128)
    vn1 = np.linalg.norm(a, axis=0)
    vn2 = vn1.copy()
    for i in range(mn):
        pvt = i + int(np.argmax(vn1[i:]))                       # idamax: first maximum
        if pvt != i:
            a[:, [pvt, i]] = a[:, [i, pvt]]
            jpvt[[pvt, i]] = jpvt[[i, pvt]]
            vn1[pvt], vn2[pvt] = vn1[i], vn2[i]
        if i < m - 1:
            beta, tau[i], a[i + 1:, i] = larfg(a[i, i], a[i + 1:, i])
        else:
            beta, tau[i], _ = larfg(a[i, i], a[i + 1:, i])
        a[i, i] = beta
        if i < n - 1:
            v = np.concatenate(([1.0], a[i + 1:, i]))
            w = v.conj() @ a[i:, i + 1:]                        # zlarf('Left', conj(tau))
            a[i:, i + 1:] -= np.conj(tau[i]) * np.outer(v, w)
        for j in range(i + 1, n):
            if vn1[j] != 0.0:
                temp = abs(a[i, j]) / vn1[j]                    # zlaqps form of the downdate
                temp = max(0.0, (1.0 + temp) * (1.0 - temp))
                temp2 = temp * (vn1[j] / vn2[j]) ** 2
                if temp2 <= TOL3Z:
                    if i < m - 1:
                        vn1[j] = np.linalg.norm(a[i + 1:, j])
                        vn2[j] = vn1[j]
                    else:
                        vn1[j] = vn2[j] = 0.0
                else:
                    vn1[j] *= np.sqrt(temp)
    return a, tau, jpvt


def laic1(job, x, sest, w, gamma):
    """zlaic1: x [j] approximate singular vector (unit norm) of a j x j triangular L with estimate sest; the
    matrix is bordered by the row [w^H gamma].  job 1: largest, job 2: smallest singular value.
    Returns (sestpr, s, c)."""
    eps = EPS
    alpha = np.vdot(x, w)
    absalp, absgam, absest = abs(alpha), abs(gamma), abs(sest)
    if job == 1:
        if sest == 0.0:
            s1 = max(absgam, absalp)
            if s1 == 0.0:
                return 0.0, 0.0, 1.0
            s, c = alpha / s1, gamma / s1
            tmp = np.sqrt((s * np.conj(s) + c * np.conj(c)).real)
            return s1 * tmp, s / tmp, c / tmp
        if absgam <= eps * absest:
            tmp = max(absest, absalp)
            s1, s2 = absest / tmp, absalp / tmp
            return tmp * np.sqrt(s1 * s1 + s2 * s2), 1.0, 0.0
        if absalp <= eps * absest:
            s1, s2 = absgam, absest
            return (s2, 1.0, 0.0) if s1 <= s2 else (s1, 0.0, 1.0)
        if absest <= eps * absalp or absest <= eps * absgam:
            s1, s2 = absgam, absalp
            if s1 <= s2:
                tmp = s1 / s2
                scl = np.sqrt(1.0 + tmp * tmp)
                return s2 * scl, (alpha / s2) / scl, (gamma / s2) / scl
            tmp = s2 / s1
            scl = np.sqrt(1.0 + tmp * tmp)
            return s1 * scl, (alpha / s1) / scl, (gamma / s1) / scl
        zeta1, zeta2 = absalp / absest, absgam / absest
        b = (1.0 - zeta1 * zeta1 - zeta2 * zeta2) * 0.5
        c = zeta1 * zeta1
        t = c / (b + np.sqrt(b * b + c)) if b > 0.0 else np.sqrt(b * b + c) - b
        sine = -(alpha / absest) / t
        cosine = -(gamma / absest) / (1.0 + t)
        tmp = np.sqrt((sine * np.conj(sine) + cosine * np.conj(cosine)).real)
        return np.sqrt(t + 1.0) * absest, sine / tmp, cosine / tmp
    # job 2
    if sest == 0.0:
        if max(absgam, absalp) == 0.0:
            sine, cosine = 1.0, 0.0
        else:
            sine, cosine = -np.conj(gamma), np.conj(alpha)
        s1 = max(abs(sine), abs(cosine))
        s, c = sine / s1, cosine / s1
        tmp = np.sqrt((s * np.conj(s) + c * np.conj(c)).real)
        return 0.0, s / tmp, c / tmp
    if absgam <= eps * absest:
        return absgam, 0.0, 1.0
    if absalp <= eps * absest:
        s1, s2 = absgam, absest
        return (s1, 0.0, 1.0) if s1 <= s2 else (s2, 1.0, 0.0)
    if absest <= eps * absalp or absest <= eps * absgam:
        s1, s2 = absgam, absalp
        if s1 <= s2:
            tmp = s1 / s2
            scl = np.sqrt(1.0 + tmp * tmp)
            return absest * (tmp / scl), -(np.conj(gamma) / s2) / scl, (np.conj(alpha) / s2) / scl
        tmp = s2 / s1
        scl = np.sqrt(1.0 + tmp * tmp)
        return absest / scl, -(np.conj(gamma) / s1) / scl, (np.conj(alpha) / s1) / scl
    zeta1, zeta2 = absalp / absest, absgam / absest
    norma = max(1.0 + zeta1 * zeta1 + zeta1 * zeta2, zeta1 * zeta2 + zeta2 * zeta2)
    test = 1.0 + 2.0 * (zeta1 - zeta2) * (zeta1 + zeta2)
    if test >= 0.0:
        b = (zeta1 * zeta1 + zeta2 * zeta2 + 1.0) * 0.5
        c = zeta2 * zeta2
        t = c / (b + np.sqrt(abs(b * b - c)))
        sine = (alpha / absest) / (1.0 - t)
        cosine = -(gamma / absest) / t
        sestpr = np.sqrt(t + 4.0 * eps * eps * norma) * absest
    else:
        b = (zeta2 * zeta2 + zeta1 * zeta1 - 1.0) * 0.5
        c = zeta1 * zeta1
        t = -c / (b + np.sqrt(b * b + c)) if b >= 0.0 else b - np.sqrt(b * b + c)
        sine = -(alpha / absest) / t
        cosine = -(gamma / absest) / (1.0 + t)
        sestpr = np.sqrt(1.0 + t + 4.0 * eps * eps * norma) * absest
    tmp = np.sqrt((sine * np.conj(sine) + cosine * np.conj(cosine)).real)
    return sestpr, sine / tmp, cosine / tmp


def gelsy_rank(r, rcond=RCOND):
    """The rank loop of zgelsy on the triangular factor r [mn, n] (upper)."""
    mn = min(r.shape)
    smax = abs(r[0, 0])
    smin = smax
    if smax == 0.0:
        return 0
    xmin = np.zeros(mn, dtype=np.complex128)
    xmax = np.zeros(mn, dtype=np.complex128)
    xmin[0] = xmax[0] = 1.0
    rank = 1
    while rank < mn:
        i = rank
        sminpr, s1, c1 = laic1(2, xmin[:rank], smin, r[:rank, i], r[i, i])
        smaxpr, s2, c2 = laic1(1, xmax[:rank], smax, r[:rank, i], r[i, i])
        if smaxpr * rcond <= sminpr:
            xmin[:rank] *= s1
            xmax[:rank] *= s2
            xmin[rank], xmax[rank] = c1, c2
            smin, smax = sminpr, smaxpr
            rank += 1
        else:
            break
    return rank


def rz(r1):
    """ztzrzf on the r x n upper-trapezoidal r1 = [R11 R12]: returns (t11 [r,r] upper, z1 [r,n]) with
    r1 = t11 @ z1 and z1 z1^H = I.  Householder reflectors applied from the right, last row first (zlatrz):
    reflector i acts on the columns {i} U {r..n-1} and annihilates row i's entries in columns r..n-1."""
    r1 = np.array(r1, dtype=np.complex128)
    r, n = r1.shape
    z1 = np.zeros((r, n), dtype=np.complex128)
    z1[np.arange(r), np.arange(r)] = 1.0
    if n == r:
        return np.triu(r1), z1
    t = r1.copy()
    g = np.eye(n, dtype=np.complex128)                           # G = G_{r-1} ... G_0,  [T 0] = r1 G
    for i in range(r - 1, -1, -1):
        cols = np.concatenate(([i], np.arange(r, n)))
        # u = t[i, cols];  larfg on u^H gives H = I - tau v v^H with u H = [beta, 0, ..., 0]
        beta, tau, v = larfg(np.conj(t[i, i]), np.conj(t[i, r:]))
        vv = np.concatenate(([1.0], v))
        blk = t[:i + 1][:, cols]
        t[:i + 1, cols] = blk - np.outer(tau * (blk @ vv), np.conj(vv))
        blk = g[:, cols]
        g[:, cols] = blk - np.outer(tau * (blk @ vv), np.conj(vv))
    return np.triu(t[:, :r]), g.conj().T[:r, :]


def gelsy_factor(a, rcond=RCOND):
    """Factored pseudo-inverse of zgelsy: x = e @ solve_triangular(t11, q1^H b).
    q1 [n, rank] orthonormal columns, t11 [rank, rank] upper, e = P Z1^H [n, rank] orthonormal columns."""
    qr, tau, jpvt = qrcp(a)
    m, n = qr.shape
    mn = min(m, n)
    r = np.triu(qr[:mn, :])
    rank = gelsy_rank(r, rcond)
    q = np.eye(m, dtype=np.complex128)
    for i in range(mn - 1, -1, -1):                             # zungqr: Q = H_0 H_1 ... H_{mn-1}
        v = np.concatenate(([1.0], qr[i + 1:, i]))
        q[i:, :] -= tau[i] * np.outer(v, v.conj() @ q[i:, :])
    t11, z1 = rz(r[:rank, :])
    e = np.zeros((n, rank), dtype=np.complex128)
    e[jpvt, :] = z1.conj().T
    return dict(q1=q[:, :rank], t11=t11, e=e, rank=rank, jpvt=jpvt, r=r)


def gelsy(a, b, rcond=RCOND):
    """x = argmin ||a x - b||, minimum norm, LAPACK zgelsy semantics.  Returns (x, rank)."""
    import scipy.linalg
    f = gelsy_factor(a, rcond)
    c = f["q1"].conj().T @ b
    c = scipy.linalg.solve_triangular(f["t11"], c)
    return f["e"] @ c, f["rank"]


def pchol_minnorm(a, b, mode):
    """Design-study variants: eps-rule pivoted Cholesky, then the minimum-norm solution of the truncated system."""
    import scipy.linalg
    x, rank, u, piv = pchol_basic(a, b, "eps", full=True)
    uu = np.zeros((rank, a.shape[0]), dtype=np.complex128)
    uu[:, :] = u[:rank]                                   # rows of U in original column order
    qt, rt = np.linalg.qr(uu.conj().T)                    # U^H = Q R
    if mode == "proj":
        return qt @ (qt.conj().T @ x), rank
    w = scipy.linalg.solve_triangular(rt, qt.conj().T @ b)
    return qt @ scipy.linalg.solve_triangular(rt, w, trans="C"), rank


def pchol_basic(a, b, rule="pstrf", full=False):
    """Round-1 GPU algorithm for comparison: diagonally pivoted Cholesky, truncated, basic solution."""
    import scipy.linalg
    a = np.array(a, dtype=np.complex128)
    n = a.shape[0]
    d = a.diagonal().real.copy()
    piv = np.arange(n)
    u = np.zeros((n, n), dtype=np.complex128)
    dmax0 = d.max()
    tol = n * 2.2e-16 * dmax0 if rule == "pstrf" else 2.2e-16 * dmax0
    rank = 0
    work = a.copy()
    for k in range(n):
        p = k + int(np.argmax(d[piv[k:]]))
        piv[[k, p]] = piv[[p, k]]
        pk = piv[k]
        if d[pk] <= tol:
            break
        ukk = np.sqrt(d[pk])
        row = (work[pk, piv[k:]] - u[:k, pk].conj() @ u[:k][:, piv[k:]]) / ukk
        u[k, piv[k:]] = row
        u[k, pk] = ukk
        d[piv[k + 1:]] -= np.abs(row[1:]) ** 2
        rank += 1
    kept = piv[:rank]
    u11 = u[:rank][:, kept]
    z = scipy.linalg.solve_triangular(u11, b[kept], trans="C")
    x = np.zeros_like(b)
    x[kept] = scipy.linalg.solve_triangular(u11, z)
    if full:
        return x, rank, u, piv
    return x, rank
