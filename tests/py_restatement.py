"""A second, independent restatement of SPEC-PROVISIONAL.md in pure Python floats (IEEE binary64, no
FMA) — small cases only.  It exists to cross-check the C oracle bit for bit; it is test
infrastructure, like oracle/.  PARITY UNPINNED w.r.t. NetTracer (no reference source exists)."""
import math


def dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def scale(a, s):
    return (a[0] * s, a[1] * s, a[2] * s)


class PyTracer:
    def __init__(self, arrays, ambient, background, eps=1e-6, max_depth=1, rules=0):
        self.rules = rules  # NT_RULE_* bits (SPEC section 8): 2 truncate, 4 inverse-square lights, 8 corner samples, 16 re-normalise
        self.a = {k: v.tolist() for k, v in arrays.items()}
        self.ambient, self.background, self.eps, self.max_depth = list(ambient), list(background), eps, max_depth
        self.ns, self.np_, self.nt = len(self.a["spheres"]), len(self.a["planes"]), len(self.a["triangles"])
        self.tri = []
        for t in self.a["triangles"]:
            v0, v1, v2 = tuple(t[0:3]), tuple(t[3:6]), tuple(t[6:9])
            e1, e2 = sub(v1, v0), sub(v2, v0)
            c = cross(e1, e2)
            self.tri.append((v0, e1, e2, scale(c, 1 / math.sqrt(dot(c, c)))))
        self.rays = 0

    def hit(self, g, o, d):
        eps = self.eps
        if g < self.ns:
            s = self.a["spheres"][g]
            oc = sub(o, s[:3])
            b = dot(oc, d)
            cc = dot(oc, oc) - s[3] * s[3]
            disc = b * b - cc
            if disc < 0:
                return None
            sq = math.sqrt(disc)
            t = -b - sq
            if not t > eps:
                t = -b + sq
            return t if t > eps else None
        g -= self.ns
        if g < self.np_:
            p = self.a["planes"][g]
            dn = dot(p[:3], d)
            if dn == 0:
                return None
            t = (p[3] - dot(p[:3], o)) / dn
            return t if t > eps else None
        g -= self.np_
        v0, e1, e2, _ = self.tri[g]
        p = cross(d, e2)
        det = dot(e1, p)
        if det == 0:
            return None
        inv = 1 / det
        tv = sub(o, v0)
        u = dot(tv, p) * inv
        if u < 0 or u > 1:
            return None
        q = cross(tv, e1)
        v = dot(d, q) * inv
        if v < 0 or u + v > 1:
            return None
        t = dot(e2, q) * inv
        return t if t > eps else None

    def nearest(self, o, d):
        best, tb = -1, math.inf
        for g in range(self.ns + self.np_ + self.nt):
            t = self.hit(g, o, d)
            if t is not None and t < tb:
                best, tb = g, t
        return best, tb

    def occluded(self, o, d, dist):
        for g in range(self.ns + self.np_ + self.nt):
            t = self.hit(g, o, d)
            if t is not None and t < dist:
                return True
        return False

    def trace(self, o, d, W, depth, acc):
        self.rays += 1
        g, t = self.nearest(o, d)
        if g < 0:
            for c in range(3):
                acc[c] = acc[c] + W * self.background[c]
            return
        P = (o[0] + d[0] * t, o[1] + d[1] * t, o[2] + d[2] * t)
        if g < self.ns:
            s = self.a["spheres"][g]
            Ng = scale(sub(P, s[:3]), 1 / s[3])
            mat = self.a["sphere_mat"][g]
        elif g < self.ns + self.np_:
            Ng = tuple(self.a["planes"][g - self.ns][:3])
            mat = self.a["plane_mat"][g - self.ns]
        else:
            Ng = self.tri[g - self.ns - self.np_][3]
            mat = self.a["triangle_mat"][g - self.ns - self.np_]
        m = self.a["materials"][mat]
        col, ka, kd, ks, shin, kr, kt, ior = m[0:3], m[3], m[4], m[5], m[6], m[7], m[8], m[9]
        entering = dot(d, Ng) < 0
        N = Ng if entering else (-Ng[0], -Ng[1], -Ng[2])
        local = [self.ambient[c] * (ka * col[c]) for c in range(3)]
        for lp in self.a["lights"]:
            Lv = sub(lp[:3], P)
            dist = math.sqrt(dot(Lv, Lv))
            L = scale(Lv, 1 / dist)
            ndl = dot(N, L)
            if not ndl > 0:
                continue
            self.rays += 1
            if self.occluded(P, L, dist):
                continue
            kdn = kd * ndl
            lc = list(lp[3:6])
            if self.rules & 4:
                att = 1 / dot(Lv, Lv)
                lc = [v * att for v in lc]
            for c in range(3):
                local[c] = local[c] + lc[c] * (col[c] * kdn)
            two = 2 * ndl
            R = (N[0] * two - L[0], N[1] * two - L[1], N[2] * two - L[2])
            rv = -dot(R, d)
            if ks > 0 and rv > 0:
                s = ks * math.pow(rv, shin)
                for c in range(3):
                    local[c] = local[c] + lc[c] * s
        for c in range(3):
            acc[c] = acc[c] + W * local[c]
        if not depth < self.max_depth:
            return
        cosi = -dot(d, N)
        wr, wt, T = kr, 0.0, None
        if kt > 0:
            eta = (1 / ior) if entering else ior
            k = 1 - (eta * eta) * (1 - cosi * cosi)
            if k < 0:
                wr = kr + kt
            else:
                wt = kt
                s = eta * cosi - math.sqrt(k)
                T = (d[0] * eta + N[0] * s, d[1] * eta + N[1] * s, d[2] * eta + N[2] * s)
        if wr > 0:
            two = 2 * cosi
            Rd = (d[0] + N[0] * two, d[1] + N[1] * two, d[2] + N[2] * two)
            if self.rules & 16:
                Rd = scale(Rd, 1 / math.sqrt(dot(Rd, Rd)))
            self.trace(P, Rd, W * wr, depth + 1, acc)
        if wt > 0:
            if self.rules & 16:
                T = scale(T, 1 / math.sqrt(dot(T, T)))
            self.trace(P, T, W * wt, depth + 1, acc)

    def render(self, cam, width, height, spp):
        n = int(round(math.sqrt(spp)))
        eye, p00, dx, dy = (list(getattr(cam, k)) for k in ("eye", "p00", "dx", "dy"))
        img = [[None] * width for _ in range(height)]
        for y in range(height):
            for x in range(width):
                tot = [0.0, 0.0, 0.0]
                for s in range(spp):
                    i, j = s % n, s // n
                    half = 0.0 if self.rules & 8 else 0.5
                    fx, fy = x + (i + half) / n, y + (j + half) / n
                    D = tuple((p00[c] + dx[c] * fx) + dy[c] * fy for c in range(3))
                    d = scale(D, 1 / math.sqrt(dot(D, D)))
                    acc = [0.0, 0.0, 0.0]
                    self.trace(tuple(eye), d, 1.0, 1, acc)
                    for c in range(3):
                        tot[c] = tot[c] + acc[c]
                px = []
                for c in range(3):
                    cv = tot[c] * (1.0 / spp)
                    px.append(0 if cv <= 0 else 255 if cv >= 1 else int(cv * 255) if self.rules & 2 else int(cv * 255 + 0.5))
                img[y][x] = px + [255]
        return img
