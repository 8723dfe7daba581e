"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Python restatement of the reference ORB extractor that calls OpenCV (cv2) for
exactly the primitives the reference calls: cv::FAST per 30-px cell, cv::resize,
cv::GaussianBlur, cv::fastAtan2 (reference src/orb_extractor.cpp).  It is the
closest executable thing to the reference in the build container (the reference
itself cannot be compiled there: OpenCV C++ headers, Eigen, g2o, FLANN are absent)
and is what pins the dependency-free C oracle (oracle/orb_oracle.c) and the golden
fixtures under tests/golden/ (tools/gen_golden.py).

Parity status: **unpinned by the reference's own tests** -- the reference has no
tests or golden vectors.  This file + cv2 4.13 (IPP/optimised
paths on or off: results identical) is the declared normative oracle.

Canonical choices where the reference is ambiguous:
  T1  quadtree tie rule: reference sorts pair<int, ExtractorNode*> i.e. by heap
      address (src/orb_extractor.cpp:684); here the pointer is replaced by the
      node's creation sequence number.
  T2  float32 ops rounded individually (no FMA); cvRound = round-half-even;
      cos/sin = float(cos(double(angle))).
  T5  a pyramid level whose FAST window is < 30 px wide or high makes the reference
      divide by zero (src/orb_extractor.cpp:784-787); canonical result = no keypoints
      on that level.
"""
from __future__ import annotations

import math

import numpy as np

try:  # cv2 exists in the build image; fixtures cover boxes where it does not
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

F32 = np.float32
PATCH_SIZE = 31
HALF_PATCH_SIZE = 15
EDGE_THRESHOLD = 19


def _load_pattern():
    import os
    txt = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "orb_pattern_data.inc")).read()
    txt = txt[txt.index("*/") + 2:]
    v = [int(t) for t in txt.replace("\n", "").split(",") if t.strip()]
    assert len(v) == 1024
    return np.array(v, dtype=np.int32).reshape(512, 2)


PATTERN = _load_pattern()


def cv_round(x) -> int:
    """cvRound: round half to even (SSE cvtss2si under the default mode)."""
    return int(np.rint(np.float64(x)))


class ExtractorCv2:
    """ORBextractor (reference include/orb_extractor.h:45-133)."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        # ctor tables: src/orb_extractor.cpp:410-470
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sf = np.float64(F32(scale_factor))  # member is double, initialised from the float arg
        self.scale = [F32(1.0)]
        self.sigma2 = [F32(1.0)]
        for i in range(1, nlevels):
            self.scale.append(F32(np.float64(self.scale[i - 1]) * sf))
            self.sigma2.append(F32(self.scale[i] * self.scale[i]))
        self.inv_scale = [F32(F32(1.0) / s) for s in self.scale]
        self.inv_sigma2 = [F32(F32(1.0) / s) for s in self.sigma2]
        factor = F32(np.float64(1.0) / sf)
        denom = F32(F32(1) - F32(math.pow(float(factor), float(nlevels))))
        nd = F32(F32(F32(nfeatures) * F32(F32(1) - factor)) / denom)
        self.per_level = []
        tot = 0
        for _ in range(nlevels - 1):
            n = cv_round(nd)
            self.per_level.append(n)
            tot += n
            nd = F32(nd * factor)
        self.per_level.append(max(nfeatures - tot, 0))
        # umax: src/orb_extractor.cpp:452-469
        umax = [0] * (HALF_PATCH_SIZE + 1)
        vmax = int(math.floor(float(F32(F32(HALF_PATCH_SIZE * F32(math.sqrt(F32(2.0)))) / F32(2)) + F32(1))))
        vmin = int(math.ceil(float(F32(F32(HALF_PATCH_SIZE * F32(math.sqrt(F32(2.0)))) / F32(2)))))
        hp2 = float(HALF_PATCH_SIZE * HALF_PATCH_SIZE)
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(hp2 - v * v))
        v0 = 0
        for v in range(HALF_PATCH_SIZE, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax
        if cv2 is not None:
            self._fast_ini = cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            self._fast_min = cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)

    # -- ComputePyramid: src/orb_extractor.cpp:1107-1132 (ring of 19 px omitted: nothing reads it)
    def level_sizes(self, w, h):
        return [(cv_round(F32(F32(w) * s)), cv_round(F32(F32(h) * s))) for s in self.inv_scale]

    def pyramid(self, img):
        h, w = img.shape
        sizes = self.level_sizes(w, h)
        pyr = [img]
        for l in range(1, self.nlevels):
            pyr.append(cv2.resize(pyr[l - 1], sizes[l], interpolation=cv2.INTER_LINEAR))
        return pyr

    # -- per-cell FAST: src/orb_extractor.cpp:765-829
    def fast_candidates(self, lvl_img):
        rows, cols = lvl_img.shape
        min_bx = min_by = EDGE_THRESHOLD - 3
        max_bx = cols - EDGE_THRESHOLD + 3
        max_by = rows - EDGE_THRESHOLD + 3
        width = F32(max_bx - min_bx)
        height = F32(max_by - min_by)
        n_cols = int(width / F32(30))
        n_rows = int(height / F32(30))
        if n_cols < 1 or n_rows < 1:  # T5: reference divides by zero here (UB); canonical = no keypoints
            return [], (min_bx, max_bx, min_by, max_by)
        w_cell = int(math.ceil(F32(width / F32(n_cols))))
        h_cell = int(math.ceil(F32(height / F32(n_rows))))
        out = []
        for i in range(n_rows):
            ini_y = min_by + i * h_cell
            max_y = ini_y + h_cell + 6
            if ini_y >= max_by - 3:
                continue
            max_y = min(max_y, max_by)
            for j in range(n_cols):
                ini_x = min_bx + j * w_cell
                max_x = ini_x + w_cell + 6
                if ini_x >= max_bx - 6:
                    continue
                max_x = min(max_x, max_bx)
                cell = np.ascontiguousarray(lvl_img[ini_y:max_y, ini_x:max_x])
                kps = self._fast_ini.detect(cell)
                if len(kps) == 0:
                    kps = self._fast_min.detect(cell)
                for k in kps:
                    out.append((F32(k.pt[0] + j * w_cell), F32(k.pt[1] + i * h_cell), F32(k.response)))
        return out, (min_bx, max_bx, min_by, max_by)

    # -- DistributeOctTree + DivideNode: src/orb_extractor.cpp:481-763, canonical tie rule T1
    @staticmethod
    def distribute(cands, min_x, max_x, min_y, max_y, n_want):
        class Node:
            __slots__ = ("x0", "x1", "y0", "y1", "keys", "nomore", "seq")
        seq_counter = [0]

        def mk(x0, x1, y0, y1):
            nd = Node()
            nd.x0, nd.x1, nd.y0, nd.y1 = x0, x1, y0, y1
            nd.keys, nd.nomore = [], False
            nd.seq = seq_counter[0]
            seq_counter[0] += 1
            return nd

        def divide(p):
            hx = int(math.ceil(F32(F32(p.x1 - p.x0) / F32(2))))
            hy = int(math.ceil(F32(F32(p.y1 - p.y0) / F32(2))))
            mx, my = p.x0 + hx, p.y0 + hy
            n1, n2 = mk(p.x0, mx, p.y0, my), mk(mx, p.x1, p.y0, my)
            n3, n4 = mk(p.x0, mx, my, p.y1), mk(mx, p.x1, my, p.y1)
            for kp in p.keys:
                if kp[0] < mx:
                    (n1 if kp[1] < my else n3).keys.append(kp)
                elif kp[1] < my:
                    n2.keys.append(kp)
                else:
                    n4.keys.append(kp)
            for c in (n1, n2, n3, n4):
                if len(c.keys) == 1:
                    c.nomore = True
            return n1, n2, n3, n4

        W, H = max_x - min_x, max_y - min_y
        n_ini = int(math.floor(float(F32(W) / F32(H)) + 0.5))  # round(): half away from zero, arg > 0
        if n_ini < 1:
            raise ValueError("image too tall: reference divides by nIni == 0")
        hx = F32(F32(W) / F32(n_ini))
        nodes = []  # python list as the std::list, index 0 = front
        roots = []
        for i in range(n_ini):
            r = mk(int(F32(hx * F32(i))), int(F32(hx * F32(i + 1))), 0, H)
            nodes.append(r)
            roots.append(r)
        for kp in cands:
            roots[int(F32(kp[0] / hx))].keys.append(kp)
        kept = []
        for nd in nodes:
            if len(nd.keys) == 1:
                nd.nomore = True
                kept.append(nd)
            elif len(nd.keys) > 1:
                kept.append(nd)
        nodes = kept
        finish = False
        while not finish:
            prev = len(nodes)
            expand = []
            new_front = []  # children in push order; final front = reversed
            survivors = []
            for nd in nodes:
                if nd.nomore:
                    survivors.append(nd)
                    continue
                for c in divide(nd):
                    if len(c.keys) > 0:
                        new_front.append(c)
                        if len(c.keys) > 1:
                            expand.append(c)
            nodes = new_front[::-1] + survivors
            if len(nodes) >= n_want or len(nodes) == prev:
                finish = True
            elif len(nodes) + 3 * len(expand) > n_want:
                while not finish:
                    prev = len(nodes)
                    order = sorted(expand, key=lambda c: (len(c.keys), c.seq))
                    expand = []
                    for nd in reversed(order):
                        front = []
                        for c in divide(nd):
                            if len(c.keys) > 0:
                                front.append(c)
                                if len(c.keys) > 1:
                                    expand.append(c)
                        nodes.remove(nd)
                        nodes = front[::-1] + nodes
                        if len(nodes) >= n_want:
                            break
                    if len(nodes) >= n_want or len(nodes) == prev:
                        finish = True
        res = []
        for nd in nodes:
            best = nd.keys[0]
            for kp in nd.keys[1:]:
                if kp[2] > best[2]:
                    best = kp
            res.append(best)
        return res

    # -- IC_Angle: src/orb_extractor.cpp:77-104
    def ic_angle(self, img, x, y):
        m01 = m10 = 0
        for u in range(-HALF_PATCH_SIZE, HALF_PATCH_SIZE + 1):
            m10 += u * int(img[y, x + u])
        for v in range(1, HALF_PATCH_SIZE + 1):
            vs = 0
            d = self.umax[v]
            for u in range(-d, d + 1):
                p, m = int(img[y + v, x + u]), int(img[y - v, x + u])
                vs += p - m
                m10 += u * (p + m)
            m01 += v * vs
        return F32(cv2.fastAtan2(float(m01), float(m10))), m01, m10

    # -- computeOrbDescriptor: src/orb_extractor.cpp:107-147
    @staticmethod
    def descriptor(blur, x, y, angle_deg):
        factor_pi = F32(np.float64(math.pi) / np.float64(F32(180.0)))
        ang = F32(F32(angle_deg) * factor_pi)
        a = F32(math.cos(float(ang)))
        b = F32(math.sin(float(ang)))
        d = np.zeros(32, dtype=np.uint8)
        for i in range(32):
            val = 0
            for k in range(8):
                t = []
                for s in (0, 1):
                    px, py = PATTERN[16 * i + 2 * k + s]
                    ry = cv_round(F32(F32(F32(px) * b) + F32(F32(py) * a)))
                    rx = cv_round(F32(F32(F32(px) * a) - F32(F32(py) * b)))
                    t.append(int(blur[y + ry, x + rx]))
                val |= (1 if t[0] < t[1] else 0) << k
            d[i] = val
        return d

    # -- extract: src/orb_extractor.cpp:1043-1105
    def extract(self, img, stages=None):
        """Returns (kps float32 n x 6 [x, y, size, angle, response, octave], desc u8 n x 32)."""
        if img.size == 0:
            return np.zeros((0, 6), F32), np.zeros((0, 32), np.uint8)
        pyr = self.pyramid(img)
        kps_all, desc_all = [], []
        if stages is not None:
            stages["pyramid"] = pyr
            stages["cands"], stages["dist"], stages["blur"] = [], [], []
        for l in range(self.nlevels):
            cands, (min_bx, max_bx, min_by, max_by) = self.fast_candidates(pyr[l])
            kept = self.distribute(cands, min_bx, max_bx, min_by, max_by, self.per_level[l]) if cands else []
            size = F32(int(F32(F32(PATCH_SIZE) * self.scale[l])))
            blur = cv2.GaussianBlur(pyr[l].copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
            if stages is not None:
                stages["cands"].append(np.array(cands, F32).reshape(-1, 3))
                stages["dist"].append(np.array(kept, F32).reshape(-1, 3))
                stages["blur"].append(blur)
            for (cx, cy, resp) in kept:
                x, y = int(cx) + min_bx, int(cy) + min_by
                ang, _, _ = self.ic_angle(pyr[l], x, y)
                desc_all.append(self.descriptor(blur, x, y, ang))
                fx, fy = F32(x), F32(y)
                if l != 0:
                    fx, fy = F32(fx * self.scale[l]), F32(fy * self.scale[l])
                kps_all.append((fx, fy, size, ang, resp, F32(l)))
        kps = np.array(kps_all, F32).reshape(-1, 6)
        desc = np.array(desc_all, np.uint8).reshape(-1, 32)
        return kps, desc
