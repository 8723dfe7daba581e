#!/usr/bin/env python3
"""Executable model of the GPU quadtree-distribution algorithm (array form).

The CUDA kernel (slam-toolkit_b200/csrc/sfe_extract.cu: octree_distribute_kernel)
is a transcription of THIS formulation: nodes are contiguous segments of a
permutation array, the std::list of the reference (src/orb_extractor.cpp:539-763)
becomes an index array rebuilt per pass by prefix sums, and candidate order is
replaced by an explicit order key.  tests/test_octree_model.py checks it against
the literal std::list emulation in the C oracle, so the reformulation is verified
on CPU before it ever runs on a GPU.
"""
import numpy as np


def order_key(x, y, w_cell, h_cell):
    """position of a candidate in the reference's vToDistributeKeys order:
    cell-row-major, raster inside the cell (src/orb_extractor.cpp:789-828)."""
    ci, cj = (y - 3) // h_cell, (x - 3) // w_cell
    return ((ci * 4096 + cj) * 4096 + y) * 4096 + x


def distribute(xs, ys, resp, W, H, n_want, w_cell, h_cell, fast_forward=True):
    """xs, ys: window-relative integer coords; returns indices of kept candidates in list order."""
    n = len(xs)
    if n == 0:
        return []
    f32 = np.float32
    n_ini = int(np.floor(f32(W) / f32(H) + f32(0.5)))
    hx = f32(f32(W) / f32(n_ini))
    # Fast-forward: with 4 * n_ini * 4^d <= n_want the first d full passes cannot end the loop through the size tests
    # (|list| <= n_ini * 4^d, |list| + 3 |E| <= 4 |list|), and node boundaries do not depend on the data, so the list after
    # those d passes is written directly: its nodes are the non-empty cells of the depth-d grid (a cell holding one point
    # stopped splitting at the depth where it became single), in the order the push_front passes produce.
    d0 = 0
    while fast_forward and 4 * n_ini * 4 ** (d0 + 1) <= n_want:
        d0 += 1
    root = [int(f32(f32(xs[p]) / hx)) for p in range(n)]

    def descend(r, x, y, depth):
        """cell index (root, q1..q_depth in base 4) and bounds of the depth-`depth` cell holding (x, y)"""
        x0, x1, y0, y1 = int(f32(hx * f32(r))), int(f32(hx * f32(r + 1))), 0, H
        cell = r
        for _ in range(depth):
            mx, my = x0 + ((x1 - x0 + 1) >> 1), y0 + ((y1 - y0 + 1) >> 1)
            q = (0 if x < mx else 1) + (0 if y < my else 2)
            x0, x1 = (x0, mx) if x < mx else (mx, x1)
            y0, y1 = (y0, my) if y < my else (my, y1)
            cell = cell * 4 + q
        return cell, (x0, x1, y0, y1)

    def cell_bounds(cell, depth):
        digits = []
        for _ in range(depth):
            digits.append(cell & 3)
            cell >>= 2
        x0, x1, y0, y1 = int(f32(hx * f32(cell))), int(f32(hx * f32(cell + 1))), 0, H
        for q in reversed(digits):
            mx, my = x0 + ((x1 - x0 + 1) >> 1), y0 + ((y1 - y0 + 1) >> 1)
            x0, x1 = (mx, x1) if q & 1 else (x0, mx)
            y0, y1 = (my, y1) if q & 2 else (y0, my)
        return x0, x1, y0, y1

    def list_order_to_cell(e, depth):
        """e-th cell of the depth-`depth` grid in list order: digit i (root = 0) runs descending when the list was reversed
        an odd number of times since that digit was appended -- root: depth odd; q_i: depth - i even."""
        digits = []
        for _ in range(depth):
            digits.append(e & 3)
            e >>= 2
        r = (n_ini - 1 - e) if depth & 1 else e
        cell = r
        for k, q in enumerate(reversed(digits)):  # k = 0 is q_1
            i = k + 1
            cell = cell * 4 + ((3 - q) if (depth - i) % 2 == 0 else q)
        return cell

    finished = False
    if d0 > 0:
        fine = [descend(root[p], int(xs[p]), int(ys[p]), d0)[0] for p in range(n)]
        cnt = [None] * (d0 + 1)
        cnt[d0] = [0] * (n_ini * 4 ** d0)
        for c in fine:
            cnt[d0][c] += 1
        for f in range(d0 - 1, -1, -1):
            cnt[f] = [sum(cnt[f + 1][4 * c:4 * c + 4]) for c in range(n_ini * 4 ** f)]
        size = [sum(1 for c in cnt[f] if c > 0) for f in range(d0 + 1)]
        d_build = d0
        for f in range(1, d0 + 1):
            if size[f] == size[f - 1]:  # a pass that did not grow the list ends the reference's loop (:665)
                d_build, finished = f, True
                break
        nodes, nodepos = [], {}
        start = 0
        for f in range(d_build, -1, -1):
            for e in range(n_ini * 4 ** f):
                c = list_order_to_cell(e, f)
                parent = cnt[f - 1][c >> 2] if f > 0 else 2
                here = cnt[f][c]
                is_node = parent > 1 and (here > 0 if f == d_build else here == 1)
                if is_node:
                    nodepos[(f, c)] = len(nodes)
                    nodes.append([*cell_bounds(c, f), start, here, 0])
                    start += here
        fill = [0] * len(nodes)
        perm, owner = [0] * n, [0] * n
        for p in range(n):
            for f in range(d_build + 1):
                c = fine[p] >> (2 * (d0 - f))
                if cnt[f][c] == 1 or f == d_build:
                    break
            li = nodepos[(f, c)]
            perm[nodes[li][4] + fill[li]] = p
            owner[nodes[li][4] + fill[li]] = li
            fill[li] += 1
    else:
        cnt = [0] * n_ini
        for r in root:
            cnt[r] += 1
        nodes = []  # list order; node = [x0,x1,y0,y1,start,cnt,eidx]
        start = 0
        starts = []
        for i in range(n_ini):
            starts.append(start)
            if cnt[i] > 0:
                nodes.append([int(f32(hx * f32(i))), int(f32(hx * f32(i + 1))), 0, H, start, cnt[i], 0])
            start += cnt[i]
        fill = [0] * n_ini
        perm = [0] * n
        for p in range(n):
            perm[starts[root[p]] + fill[root[p]]] = p
            fill[root[p]] += 1
        owner = [0] * n
        for li, nd in enumerate(nodes):
            for p in range(nd[4], nd[4] + nd[5]):
                owner[p] = li
        e = 0
        for nd in nodes:  # E order for roots is irrelevant (first pass is always a full pass)
            if nd[5] > 1:
                nd[6] = e
                e += 1

    def split_pass(careful):
        nonlocal nodes, perm, owner
        nL = len(nodes)
        childcnt = [[0, 0, 0, 0] for _ in range(nL)]
        qs = [None] * n
        for p in range(n):
            i = owner[p]
            x0, x1, y0, y1, st, c, _ = nodes[i]
            if c > 1:
                mx, my = x0 + ((x1 - x0 + 1) >> 1), y0 + ((y1 - y0 + 1) >> 1)
                q = (0 if xs[perm[p]] < mx else 1) + (0 if ys[perm[p]] < my else 2)
                qs[p] = (q, childcnt[i][q])
                childcnt[i][q] += 1
        nz = [sum(1 for q in range(4) if childcnt[i][q] > 0) if nodes[i][5] > 1 else 0 for i in range(nL)]
        ne = [sum(1 for q in range(4) if childcnt[i][q] > 1) if nodes[i][5] > 1 else 0 for i in range(nL)]
        expandable = [i for i in range(nL) if nodes[i][5] > 1]
        if not careful:
            t_of = {i: t for t, i in enumerate(expandable)}
            n_split = len(expandable)
        else:
            keys = {i: (nodes[i][5], nodes[i][6]) for i in expandable}
            t_of = {i: sum(1 for j in expandable if keys[j] > keys[i]) for i in expandable}
            by_t = sorted(expandable, key=lambda i: t_of[i])
            size = nL
            n_split = 0
            for t, i in enumerate(by_t):
                size += nz[i] - 1
                n_split = t + 1
                if size >= n_want:
                    break
        split = [False] * nL
        for i in expandable:
            if t_of[i] < n_split:
                split[i] = True
        by_t = sorted([i for i in expandable if split[i]], key=lambda i: t_of[i])
        pushpre, epre, acc_p, acc_e = {}, {}, 0, 0
        for i in by_t:
            pushpre[i], epre[i] = acc_p, acc_e
            acc_p += nz[i]
            acc_e += ne[i]
        C = acc_p
        n_unsplit = nL - len(by_t)
        new_nodes = [None] * (C + n_unsplit)
        childpos = [[-1] * 4 for _ in range(nL)]
        ur = 0
        for i in range(nL):
            x0, x1, y0, y1, st, c, ei = nodes[i]
            if split[i]:
                mx, my = x0 + ((x1 - x0 + 1) >> 1), y0 + ((y1 - y0 + 1) >> 1)
                boxes = [(x0, mx, y0, my), (mx, x1, y0, my), (x0, mx, my, y1), (mx, x1, my, y1)]
                k = ke = 0
                off = 0
                for q in range(4):
                    cc = childcnt[i][q]
                    if cc > 0:
                        pos = C - 1 - (pushpre[i] + k)
                        new_nodes[pos] = [*boxes[q], st + off, cc, epre[i] + ke if cc > 1 else 0]
                        childpos[i][q] = pos
                        k += 1
                        if cc > 1:
                            ke += 1
                    off += cc
            else:
                pos = C + ur
                ur += 1
                new_nodes[pos] = list(nodes[i])
                childpos[i][0] = pos
        new_perm, new_owner = [0] * n, [0] * n
        for p in range(n):
            i = owner[p]
            if split[i]:
                q, s = qs[p]
                off = sum(childcnt[i][:q])
                np_ = nodes[i][4] + off + s
                new_perm[np_] = perm[p]
                new_owner[np_] = childpos[i][q]
            else:
                new_perm[p] = perm[p]
                new_owner[p] = childpos[i][0]
        nodes, perm, owner = new_nodes, new_perm, new_owner
        return acc_e

    while not finished:
        prev = len(nodes)
        n_expand = split_pass(False)
        if len(nodes) >= n_want or len(nodes) == prev:
            break
        if len(nodes) + 3 * n_expand > n_want:
            while True:
                prev = len(nodes)
                split_pass(True)
                if len(nodes) >= n_want or len(nodes) == prev:
                    break
            break
    out = []
    for nd in nodes:
        best = None
        for p in range(nd[4], nd[4] + nd[5]):
            c = perm[p]
            k = (-int(resp[c]), order_key(int(xs[c]), int(ys[c]), w_cell, h_cell))
            if best is None or k < best[0]:
                best = (k, c)
        out.append(best[1])
    return out
