"""Seeded synthetic MIPs of the shapes BASELINE.json names, as raw CSR problem dicts.

Every generator plants a feasible point ``x*`` first and derives sides and fixings from it, so the instance is
feasible by construction unless ``infeasible=True`` asks for contradicting fixings.  The same dict feeds the CUDA
path (``scip_b200.LinearPropagator``), the CPU oracle and -- written with ``lpb.write_lpb`` -- the reference
driver, which rebuilds it through ``SCIPcreateConsBasicLinear``.

Problem dict keys: rowptr[int64 nrows+1], colidx[int32 nnz], vals[f64 nnz], lhs/rhs[f64 nrows],
lb/ub[f64 ncols], vartype[u8 ncols] (0 continuous, 1 integral).
"""
from __future__ import annotations

import numpy as np

INF = 1e20


def _row_lengths(rng, nrows, minlen, maxlen, nnz):
    """lengths in [minlen,maxlen] from an exponentially tilted uniform law whose mean is nnz/nrows, then
    nudged so that they sum to exactly ``nnz``"""
    mean = nnz / nrows
    if not (minlen <= mean <= maxlen):
        raise ValueError("nnz/nrows outside [minlen,maxlen]")
    ls = np.arange(minlen, maxlen + 1, dtype=np.float64)
    lo, hi = -5.0, 5.0
    for _ in range(200):
        t = 0.5 * (lo + hi)
        w = np.exp(t * (ls - ls.mean()))
        m = float((w * ls).sum() / w.sum())
        if m < mean:
            lo = t
        else:
            hi = t
    w = np.exp(0.5 * (lo + hi) * (ls - ls.mean()))
    lens = rng.choice(ls.astype(np.int64), size=nrows, p=w / w.sum())
    diff = int(nnz - lens.sum())
    step = 1 if diff > 0 else -1
    while diff != 0:
        cand = np.flatnonzero((lens < maxlen) if step > 0 else (lens > minlen))
        take = cand[rng.permutation(len(cand))[: abs(diff)]]
        lens[take] += step
        diff = int(nnz - lens.sum())
    return lens


def _draw_columns(rng, lens, ncols):
    """uniform column indices, without repetition inside a row"""
    rowptr = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    nnz = int(rowptr[-1])
    rowid = np.repeat(np.arange(len(lens), dtype=np.int64), lens)
    cols = rng.integers(0, ncols, size=nnz, dtype=np.int64)
    for _ in range(100):
        key = rowid * ncols + cols
        order = np.argsort(key, kind="stable")
        sk = key[order]
        dup = np.zeros(nnz, dtype=bool)
        dup[order[1:]] = sk[1:] == sk[:-1]
        nd = int(dup.sum())
        if nd == 0:
            break
        cols[dup] = rng.integers(0, ncols, size=nd, dtype=np.int64)
    else:
        raise RuntimeError("could not remove duplicate columns")
    return rowptr, rowid, cols


def setcover(nrows=1_000_000, ncols=1_000_000, nnz=10_000_000, seed=1, minlen=5, maxlen=20,
             cover_frac=0.6, pack_frac=0.25, fix_frac=0.10, ones_frac=0.15, infeasible=False):
    """Config C3: binaries only; a mix of covering rows (sum x >= 1), packing rows (sum x <= 1) and knapsack rows
    (sum a x <= b, a in 1..20), consistent with a planted 0/1 point; ``fix_frac`` of the variables are fixed to
    their planted value, which starts the propagation cascade.  ``infeasible=True`` fixes 40 % of all variables
    to 0 instead (some covering row then has no free variable)."""
    rng = np.random.default_rng(seed)
    lens = _row_lengths(rng, nrows, minlen, maxlen, nnz)
    rowptr, rowid, cols = _draw_columns(rng, lens, ncols)
    xstar = (rng.random(ncols) < ones_frac)
    ones = np.flatnonzero(xstar)
    zeros = np.flatnonzero(~xstar)
    kind = rng.random(nrows)
    is_cover = kind < cover_frac
    is_pack = (kind >= cover_frac) & (kind < cover_frac + pack_frac)
    is_knap = ~(is_cover | is_pack)
    vals = np.ones(nnz, dtype=np.float64)
    first = rowptr[:-1]

    # packing rows: at most one planted 1 -> all positions but the first are redrawn among planted zeros
    pk = is_pack[rowid]
    notfirst = np.ones(nnz, dtype=bool)
    notfirst[first] = False
    redo = pk & notfirst & xstar[cols]
    cols[redo] = zeros[rng.integers(0, len(zeros), size=int(redo.sum()))]
    # covering rows: at least one planted 1 -> the first position is redrawn among planted ones if needed
    act = np.add.reduceat(xstar[cols].astype(np.int64), first)
    need = is_cover & (act == 0)
    cols[first[need]] = ones[rng.integers(0, len(ones), size=int(need.sum()))]
    # the redraws may have created repeated columns in a few rows: drop those entries
    key = rowid * ncols + cols
    order = np.argsort(key, kind="stable")
    sk = key[order]
    dup = np.zeros(nnz, dtype=bool)
    dup[order[1:]] = sk[1:] == sk[:-1]
    if dup.any():
        keep = ~dup
        cols, vals, rowid = cols[keep], vals[keep], rowid[keep]
        lens = np.bincount(rowid, minlength=nrows).astype(np.int64)
        rowptr = np.zeros(nrows + 1, dtype=np.int64)
        np.cumsum(lens, out=rowptr[1:])
        nnz = int(rowptr[-1])
        first = rowptr[:-1]
    kn = is_knap[rowid]
    vals[kn] = rng.integers(1, 21, size=int(kn.sum())).astype(np.float64)

    actx = np.add.reduceat(vals * xstar[cols], first)
    lhs = np.full(nrows, -INF)
    rhs = np.full(nrows, INF)
    lhs[is_cover] = 1.0
    rhs[is_pack] = 1.0
    slack = rng.integers(0, 15, size=nrows).astype(np.float64)
    rhs[is_knap] = actx[is_knap] + slack[is_knap]

    lb = np.zeros(ncols)
    ub = np.ones(ncols)
    if infeasible:
        fixed = rng.random(ncols) < 0.40
        ub[fixed] = 0.0
    else:
        fixed = rng.random(ncols) < fix_frac
        lb[fixed & xstar] = 1.0
        ub[fixed & ~xstar] = 0.0
    return dict(rowptr=rowptr, colidx=cols.astype(np.int32), vals=vals, lhs=lhs, rhs=rhs, lb=lb, ub=ub,
                vartype=np.ones(ncols, dtype=np.uint8), name=f"setcover_{nrows}x{ncols}_s{seed}")


def mixed_knapsack(nrows=200_000, ncols=2_000_000, nnz=50_000_000, seed=2, dense_frac=0.01,
                   len_range=(50, 350), dense_range=(2_000, 8_000), eq_frac=0.10, fix_frac=0.01,
                   offbound_frac=0.02, infeasible=False):
    """Config C4: 50 % binaries, 30 % general integers in [0,U] (U in 10/100/1000), 20 % continuous in [0,1000];
    knapsack-like rows with integer coefficients 1..100, negative on a fixed 20 % of the columns; 1 % long dense
    rows (block-per-row path); 10 % ranged/equality rows.  The planted point sits on the activity-minimising
    corner except for ``offbound_frac`` of the variables, which keeps every row tight enough to propagate."""
    rng = np.random.default_rng(seed)
    ndense = int(round(nrows * dense_frac))
    nnorm = nrows - ndense
    dense_lens = rng.integers(dense_range[0], dense_range[1] + 1, size=ndense)
    dense_lens = np.minimum(dense_lens, ncols // 2)
    rest = nnz - int(dense_lens.sum())
    lo, hi = len_range
    hi = min(hi, ncols // 2)
    rest = min(max(rest, nnorm * lo), nnorm * hi)
    norm_lens = _row_lengths(rng, nnorm, lo, hi, rest)
    lens = np.concatenate([norm_lens, dense_lens])
    lens = lens[rng.permutation(nrows)]
    rowptr, rowid, cols = _draw_columns(rng, lens, ncols)
    nnz = int(rowptr[-1])
    first = rowptr[:-1]

    t = rng.random(ncols)
    vartype = (t < 0.8).astype(np.uint8)
    ub = np.ones(ncols)
    isint = (t >= 0.5) & (t < 0.8)
    ub[isint] = rng.choice(np.array([10.0, 100.0, 1000.0]), size=int(isint.sum()))
    ub[t >= 0.8] = 1000.0
    lb = np.zeros(ncols)
    negcol = rng.random(ncols) < 0.20
    vals = rng.integers(1, 101, size=nnz).astype(np.float64)
    vals[negcol[cols]] *= -1.0

    # planted point: activity-minimising corner, a few variables moved into the interior
    xstar = np.where(negcol, ub, lb)
    off = rng.random(ncols) < offbound_frac
    frac = rng.random(ncols)
    moved = np.where(negcol, ub - frac * (ub - lb), lb + frac * (ub - lb))
    moved = np.where(vartype != 0, np.round(moved), moved)
    xstar = np.where(off, moved, xstar)

    actx = np.add.reduceat(vals * xstar[cols], first)
    alpha = np.abs(vals) * (ub - lb)[cols]
    maxalpha = np.maximum.reduceat(alpha, first)
    rhs = np.floor(actx + rng.random(nrows) * 0.5 * maxalpha) + 1.0
    lhs = np.full(nrows, -INF)
    eq = rng.random(nrows) < eq_frac
    lhs[eq] = np.ceil(actx[eq] - rng.random(int(eq.sum())) * 0.5 * maxalpha[eq]) - 1.0

    fixed = rng.random(ncols) < fix_frac
    lb = np.where(fixed, xstar, lb)
    ub = np.where(fixed, xstar, ub)
    if infeasible:
        r = int(rng.integers(0, nrows))
        rhs[r] = np.floor((vals[rowptr[r]:rowptr[r + 1]] * np.where(vals[rowptr[r]:rowptr[r + 1]] > 0,
                           lb[cols[rowptr[r]:rowptr[r + 1]]], ub[cols[rowptr[r]:rowptr[r + 1]]])).sum()) - 5.0
    return dict(rowptr=rowptr, colidx=cols.astype(np.int32), vals=vals, lhs=lhs, rhs=rhs, lb=lb, ub=ub,
                vartype=vartype, name=f"mixedknap_{nrows}x{ncols}_s{seed}")


def unit_network(nrows=200_000, ncols=150_000, nnz=1_000_000, seed=4, minlen=2, maxlen=8, fix_frac=0.15,
                 unbounded_frac=0.03):
    """Rows with coefficients +1 / -1 only (precedence, flow-balance, cardinality shapes) over binaries, small general
    integers and continuous variables, some of them without an upper bound; <=, >=, ranged and equality rows around
    a planted point.  Exercises the unit-row storage class of the filter sweep (sign flags, no values read) together
    with bound gathers, integrality rounding and infinite contributions."""
    rng = np.random.default_rng(seed)
    lens = _row_lengths(rng, nrows, minlen, maxlen, nnz)
    rowptr, rowid, cols = _draw_columns(rng, lens, ncols)
    nnz = int(rowptr[-1])
    first = rowptr[:-1]
    vals = np.where(rng.random(nnz) < 0.45, -1.0, 1.0)

    t = rng.random(ncols)
    vartype = (t < 0.85).astype(np.uint8)
    lb = np.zeros(ncols)
    ub = np.ones(ncols)
    isint = (t >= 0.6) & (t < 0.85)
    ub[isint] = rng.choice(np.array([3.0, 10.0]), size=int(isint.sum()))
    cont = t >= 0.85
    ub[cont] = 7.5
    lb[cont & (rng.random(ncols) < 0.3)] = -2.25
    xstar = lb + rng.random(ncols) * (ub - lb)
    xstar = np.where(vartype != 0, np.round(xstar), xstar)
    unb = cont & (rng.random(ncols) < unbounded_frac / 0.15)
    ub[unb] = INF

    actx = np.add.reduceat(vals * xstar[cols], first)
    kind = rng.random(nrows)
    s1 = rng.integers(0, 3, size=nrows).astype(np.float64)
    s2 = rng.integers(0, 3, size=nrows).astype(np.float64)
    lhs = np.full(nrows, -INF)
    rhs = np.full(nrows, INF)
    le = kind < 0.4
    ge = (kind >= 0.4) & (kind < 0.7)
    rg = (kind >= 0.7) & (kind < 0.9)
    eq = kind >= 0.9
    rhs[le] = actx[le] + s1[le]
    lhs[ge] = actx[ge] - s1[ge]
    lhs[rg] = actx[rg] - s1[rg]
    rhs[rg] = actx[rg] + s2[rg]
    lhs[eq] = actx[eq]
    rhs[eq] = actx[eq]

    fixed = rng.random(ncols) < fix_frac
    lb = np.where(fixed, xstar, lb)
    ub = np.where(fixed, xstar, ub)
    return dict(rowptr=rowptr, colidx=cols.astype(np.int32), vals=vals, lhs=lhs, rhs=rhs, lb=lb + 0.0, ub=ub + 0.0,
                vartype=vartype, name=f"unitnet_{nrows}x{ncols}_s{seed}")


def probing_batch(prob, nvec=1024, seed=3):
    """Config C5: ``nvec`` bound vectors = base bounds + one seeded unfixed variable fixed to 0 or 1 each (the
    SCIPapplyProbingVar pattern, prop_probing.c:1254-1279).  Returns (lb[nvec,ncols], ub[nvec,ncols], var, val)."""
    rng = np.random.default_rng(seed)
    free = np.flatnonzero(prob["lb"] < prob["ub"])
    var = free[rng.integers(0, len(free), size=nvec)]
    val = rng.integers(0, 2, size=nvec)
    lb = np.repeat(prob["lb"][None, :], nvec, axis=0)
    ub = np.repeat(prob["ub"][None, :], nvec, axis=0)
    fixval = np.where(val == 1, prob["ub"][var], prob["lb"][var])
    lb[np.arange(nvec), var] = fixval
    ub[np.arange(nvec), var] = fixval
    return lb, ub, var, val


def _from_rows(rows, lhs, rhs, lb, ub, vartype):
    rowptr = np.cumsum([0] + [len(r) for r in rows]).astype(np.int64)
    return dict(rowptr=rowptr, colidx=np.array([c for r in rows for c, _ in r], dtype=np.int32),
                vals=np.array([v for r in rows for _, v in r], dtype=np.float64),
                lhs=np.array(lhs, dtype=np.float64), rhs=np.array(rhs, dtype=np.float64),
                lb=np.array(lb, dtype=np.float64), ub=np.array(ub, dtype=np.float64),
                vartype=np.array(vartype, dtype=np.uint8))


def edge_cancellation():
    """rows whose activities cancel by twelve orders of magnitude once a singleton row has fixed a variable -- what the
    reference answers with consdataGetReliableResidualActivity (cons_linear.c:2580-2657) after an unreliable incremental
    update: 1e12 x0 - 1e12 x1 + z <= 1 with x1 <= 0 arriving from its own row, for continuous and for
    integer x / z, plus the mirrored >= row"""
    rows, lhs, rhs, lb, ub, vt = [], [], [], [], [], []
    for integral in (0, 1):
        b = len(lb)
        x0, x1, z, y0, y1, w = range(b, b + 6)
        lb += [0.0, 0.0, 0.0, 0.0, 0.0, -10.0]
        ub += [1.0, 1.0, 10.0, 1.0, 1.0, 10.0]
        vt += [integral, integral, integral, integral, integral, 0]
        rows += [[(x0, 1e12), (x1, -1e12), (z, 1.0)], [(x1, 1.0)],
                 [(y0, -1e12), (y1, 1e12), (w, 1.0)], [(y1, 3.0)]]
        lhs += [-INF, -INF, -1.0, -INF]
        rhs += [1.0, 0.0, INF, 0.0]
    return _from_rows(rows, lhs, rhs, lb, ub, vt)


def edge_huge(infeasible=False):
    """rows with three and more contributions of 1e15 or more each: they are counted, not summed (cons_linear.c:1773-1948),
    and enter the relaxed activity of the row verdict as count * hugeval (:2386, :2481) -- the verdict depends on the true
    count.  infeasible: the side sits between two and three times hugeval"""
    n = 5
    lb = [1e6] * n + [0.0, 0.0]
    ub = [2e6] * n + [5.0, 8.0]
    vt = [0] * n + [1, 0]
    rows = [[(0, 1e10), (1, 1e10), (2, 1e10), (5, 1.0)],
            [(0, -1e10), (1, -1e10), (2, -1e10), (3, -1e10), (6, 2.0)],
            [(5, 1.0), (6, 1.0)]]
    lhs = [-INF, -4.5e15 if not infeasible else -3.5e15, -INF]
    rhs = [3.5e15 if not infeasible else 2.5e15, INF, 6.0]
    return _from_rows(rows, lhs, rhs, lb, ub, vt)


def ranged_rows(nrows=400, ncols=600, seed=31, infeasible=False):
    """equations and ranged rows with a divisibility structure -- what the reference's ranged-row propagation looks at
    (rangedRowPropagation, cons_linear.c:5715-6696): integer variables whose coefficients share a divisor g >= 2, beside
    one variable (or a few) with a coefficient of 1 or one that is coprime to g.  Row kinds: (A) g-group + one odd
    variable, equation; (B) the same with sides a little apart; (C) one variable with coefficient g beside several +1 / -1
    variables; (D) with a continuous variable (only the infeasibility test applies); (E) plain inequalities that couple the
    variables.  A feasible point is planted; ``infeasible`` adds the reference's own example 12 x1 + 9 x2 - x3 = 0 with
    x3 in [1, 2]."""
    rng = np.random.default_rng(seed)
    kind = rng.random(ncols)
    vartype = (kind < 0.85).astype(np.uint8)
    isbin = kind < 0.25
    ub = np.where(isbin, 1.0, np.where(vartype != 0, rng.integers(3, 61, size=ncols).astype(np.float64), 10.0))
    lb = np.zeros(ncols)
    xstar = np.where(vartype != 0, np.floor(rng.random(ncols) * (ub + 1.0)), np.round(rng.random(ncols) * ub, 3))
    xstar = np.minimum(xstar, ub)
    fixed = rng.random(ncols) < 0.04
    lb = np.where(fixed, xstar, lb)
    ub = np.where(fixed, xstar, ub)
    ints = np.flatnonzero((vartype != 0) & ~isbin)
    bins = np.flatnonzero(isbin)
    conts = np.flatnonzero(vartype == 0)
    rows, lhs, rhs = [], [], []
    for r in range(nrows):
        t = rng.random()
        g = int(rng.choice([2, 3, 4, 5, 6, 10, 12]))
        if t < 0.35 or t >= 0.9:            # (A) / (D)
            k = int(rng.integers(2, 7))
            cols = rng.choice(ints, size=k + 1, replace=False)
            coefs = [float(g * int(rng.integers(1, 6)) * (1 if rng.random() < 0.7 else -1)) for _ in range(k)]
            odd = [1.0, -1.0, float(g + 1), float(2 * g - 1)][int(rng.integers(0, 4))]
            row = list(zip(cols[:k].tolist(), coefs)) + [(int(cols[k]), odd)]
            if t >= 0.9 and len(conts) > 0:
                row.append((int(rng.choice(conts)), float(rng.integers(1, 4))))
            act = sum(a * xstar[j] for j, a in row)
            rows.append(row)
            lhs.append(act)
            rhs.append(act)
        elif t < 0.55:                      # (B)
            k = int(rng.integers(2, 6))
            cols = rng.choice(ints, size=k + 1, replace=False)
            row = [(int(cols[i]), float(g * int(rng.integers(1, 5)))) for i in range(k)] + [(int(cols[k]), 1.0)]
            act = sum(a * xstar[j] for j, a in row)
            rows.append(row)
            lhs.append(act - float(rng.integers(0, max(g - 1, 1))))
            rhs.append(act + float(rng.integers(0, max(g - 1, 1))))
        elif t < 0.75:                      # (C)
            k = int(rng.integers(2, 6))
            z = int(rng.choice(ints))
            others = rng.choice(bins, size=k, replace=False)
            row = [(z, float(g * (1 if rng.random() < 0.8 else -1)))] + [(int(j), 1.0 if rng.random() < 0.6 else -1.0) for j in others]
            act = sum(a * xstar[j] for j, a in row)
            rows.append(row)
            lhs.append(act)
            rhs.append(act + float(rng.integers(0, 2)))
        else:                               # (E)
            k = int(rng.integers(3, 9))
            cols = rng.choice(ncols, size=k, replace=False)
            row = [(int(j), float(rng.integers(1, 9)) * (1 if rng.random() < 0.8 else -1)) for j in cols]
            act = sum(a * xstar[j] for j, a in row)
            rows.append(row)
            lhs.append(-INF)
            rhs.append(act + float(rng.integers(0, 6)))
    if infeasible:
        # three fresh integer columns: only the divisibility argument sees that the row has no solution
        x1, x2, x3 = ncols, ncols + 1, ncols + 2
        lb = np.concatenate([lb, [-INF, -INF, 1.0]])
        ub = np.concatenate([ub, [INF, INF, 2.0]])
        vartype = np.concatenate([vartype, np.ones(3, dtype=np.uint8)])
        rows.append([(x1, 12.0), (x2, 9.0), (x3, -1.0)])
        lhs.append(0.0)
        rhs.append(0.0)
    prob = _from_rows(rows, lhs, rhs, lb, ub, vartype)
    prob["name"] = f"ranged_rows_{nrows}x{ncols}_s{seed}"
    return prob
