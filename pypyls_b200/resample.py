# -*- coding: utf-8 -*-
"""
Host-side resampling tables that replay the reference's NumPy random stream.

``gen_permsamp`` / ``gen_bootsamp`` / ``gen_splits`` return, for a given seed,
exactly the tables of pyls/base.py:10-79, :82-159 and :162-229 (same draws from the same
``RandomState`` in the same order), which is what "identical seeds" parity
against the CPU reference needs.  The only change is how duplicates are found:
the reference compares each candidate column with all earlier ones
(O(n^2 * S)); here the earlier columns' bytes are kept in a set, which accepts
and rejects exactly the same candidates.

Throughput runs use the on-device generator instead
(``ResamplingEngine.gen_perm_indices`` / ``gen_boot_indices``).
"""

import warnings

import numpy as np


def check_random_state(seed):
    """int | RandomState | None -> RandomState, as sklearn's helper of the same
    name (used at pyls/base.py:37,109,283)."""
    if seed is None or seed is np.random:
        return np.random.mtrand._rand
    if isinstance(seed, (int, np.integer)):
        return np.random.RandomState(seed)
    if isinstance(seed, np.random.RandomState):
        return seed
    raise ValueError('%r cannot be used to seed a numpy.random.RandomState'
                     ' instance' % seed)


def _layout(groups, n_cond):
    """Row ids as an (n_cond, n_subj) table (column = subject, rows ordered
    group -> condition -> subject) and the subject bounds of every group."""
    groups = [int(g) for g in groups]
    bounds = np.concatenate([[0], np.cumsum(groups)]).astype(int)
    cols, row0 = [], 0
    for g in groups:
        cols.append(row0 + np.arange(n_cond * g).reshape(n_cond, g))
        row0 += n_cond * g
    return groups, bounds, cols


def _column(table, order, bounds):
    """Source row of every destination row when subject slot k takes subject
    order[k]: groups stacked, condition-major inside a group."""
    picked = table[:, order]
    return np.concatenate([picked[:, a:b].ravel()
                           for a, b in zip(bounds[:-1], bounds[1:])])


def gen_permsamp(groups, n_cond, n_perm, seed=None, verbose=True):
    """
    Permutation table (S, n_perm); bit-identical to pyls/base.py:10-79 for the
    same ``seed`` state.
    """
    groups, bounds, blocks = _layout(groups, n_cond)
    n_subj = int(bounds[-1])
    rs = check_random_state(seed)
    out = np.zeros((n_subj * n_cond, n_perm), dtype=int)
    seen = set()
    warned = False
    for i in range(n_perm):
        tries = 0
        while True:
            tries += 1
            # per-subject shuffle of the condition order (utils.permute_cols:
            # one uniform matrix per group block, arg-sorted along conditions)
            shuffled = []
            for blk in blocks:
                order = rs.random_sample(blk.shape).argsort(axis=0)
                shuffled.append(np.take_along_axis(blk, order, axis=0))
            table = np.concatenate(shuffled, axis=1)
            perm = rs.permutation(n_subj)
            bad = False
            if len(groups) > 1:
                for a, b in zip(bounds[:-1], bounds[1:]):
                    seg = perm[a:b]
                    if seg.min() >= a and seg.max() < b:
                        bad = True
            col = _column(table, perm, bounds)
            key = col.tobytes()
            if key in seen:
                bad = True
            if not bad or tries >= 500:
                break
        if tries == 500 and not warned:
            warnings.warn('WARNING: Duplicate permutations used.')
            warned = True
        out[:, i] = col
        seen.add(key)
    return out


def gen_bootsamp(groups, n_cond, n_boot, seed=None, verbose=True):
    """
    Bootstrap table (S, n_boot); bit-identical to pyls/base.py:82-159 for the
    same ``seed`` state (including the reference's per-group duplicate rule,
    which compares rows [a, b) with a, b the group's SUBJECT bounds).
    """
    groups, bounds, blocks = _layout(groups, n_cond)
    n_subj = int(bounds[-1])
    rs = check_random_state(seed)
    table = np.concatenate(blocks, axis=1)
    min_subj = int(np.ceil(min(groups) * 0.5))
    out = np.zeros((n_subj * n_cond, n_boot), dtype=int)
    seen = [set() for _ in groups]
    warned = False
    for i in range(n_boot):
        tries = 0
        while True:
            tries += 1
            boot = np.zeros(n_subj, dtype=int)
            for a, b in zip(bounds[:-1], bounds[1:]):
                pool = np.arange(a, b)
                while True:
                    draw = np.sort(rs.choice(pool, size=b - a, replace=True))
                    if np.unique(draw).size >= min_subj:
                        break
                boot[a:b] = draw
            col = _column(table, boot, bounds)
            keys = [col[a:b].tobytes()
                    for a, b in zip(bounds[:-1], bounds[1:])]
            bad = any(k in s for k, s in zip(keys, seen))
            if not bad or tries >= 500:
                break
        if tries == 500 and not warned:
            warnings.warn('WARNING: Duplicate bootstraps used.')
            warned = True
        out[:, i] = col
        for k, s in zip(keys, seen):
            s.add(k)
    return out


def gen_splits(groups, n_cond, n_split, seed=None, test_size=0.5):
    """
    Train / split masks (S, n_split) bool; bit-identical to
    pyls/base.py:162-229 for the same ``seed`` state: per group one
    ``randint(0, 2)`` (ceil or floor of n_g * (1 - test_size) subjects) and one
    ``permutation(n_g)`` whose head is the chosen subjects -- the draws
    ``RandomState.choice`` makes there.
    """
    groups, bounds, _ = _layout(groups, n_cond)
    n_subj = int(bounds[-1])
    rs = check_random_state(seed)
    out = np.zeros((n_subj * n_cond, n_split), dtype=bool)
    seen = set()
    warned = False
    for i in range(n_split):
        tries = 0
        while True:
            tries += 1
            rows = []
            for a, b in zip(bounds[:-1], bounds[1:]):
                frac = (b - a) * (1 - test_size)
                num = int(np.floor(frac) if rs.randint(0, 2) else np.ceil(frac))
                mask = np.zeros(b - a, dtype=bool)
                mask[rs.permutation(b - a)[:num]] = True
                rows.append(np.tile(mask, n_cond))
            col = np.concatenate(rows)
            key = col.tobytes()
            if key not in seen or tries >= 500:
                break
        if tries == 500 and not warned:
            warnings.warn('WARNING: Duplicate split halves used.')
            warned = True
        out[:, i] = col
        seen.add(key)
    return out


def cell_of_rows(groups, n_cond):
    """Cell (group x condition) of every row: rows are ordered group ->
    condition -> subject (pyls/structures.py:37-44)."""
    return np.repeat(np.arange(len(groups) * n_cond),
                     np.repeat([int(g) for g in groups], n_cond))


def check_bootsamples(table, groups, n_cond):
    """A bootstrap table must draw every row from its own group x condition
    cell, as gen_bootsamp does (pyls/base.py:134-143): the engine contracts
    each cell's operand rows over that cell's rows of X only, so rows drawn
    from another cell would be left out silently."""
    cells = cell_of_rows(groups, n_cond)
    table = np.asarray(table)
    if table.shape[0] != cells.size:
        return            # the shape check of the caller reports this
    if not np.array_equal(cells[table], np.broadcast_to(cells[:, None],
                                                        table.shape)):
        raise ValueError('Provided `bootsamples` draw rows from other '
                         'group / condition cells; bootstrap resamples must '
                         'keep every row inside its cell (as gen_bootsamp '
                         'does).')


def shard_range(n, rank, world_size):
    """Contiguous block [first, first + count) of ``n`` resample ids owned by
    ``rank``; blocks differ in size by at most one."""
    base, extra = divmod(int(n), int(world_size))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)
