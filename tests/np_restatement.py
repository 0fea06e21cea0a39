"""Second, independent restatement of the reference's CPU path — scalar numpy, pure-Python loops.

TEST INFRASTRUCTURE ONLY.  Written from the reference source without looking at oracle/sph_oracle.cpp's
code paths: np.float32 scalars force one fp32 rounding per operation (QVector3D semantics), Python
floats are the doubles of the reference's mixed arithmetic.  It is slow (small scenes only) and exists
to pin the C++ oracle: both must agree BIT FOR BIT (tests/test_oracle_vs_np_restatement.py).

Citations are relative to /root/reference.
"""
import math

import numpy as np

f32 = np.float32
H = f32(0.0457)  # include/CParticle.h:80
VISCOSITY, MASS, GAS_STIFFNESS, REST_DENSITY = f32(3.5), f32(0.02), f32(3.0), f32(998.29)  # :81-84
WALL_K, WALL_DAMPING = 10000.0, -0.9  # include/CCollisionGeometry.h:20-21 (double literals)
DT = f32(0.01)  # src/CBaseParticleSimulator.cpp:7


def vsub(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def vadd(a, b):
    return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]


def vmul(a, s):  # QVector3D * float: the factor is narrowed to float first
    s = f32(s)
    return [a[0] * s, a[1] * s, a[2] * s]


def vdiv(a, s):
    s = f32(s)
    return [a[0] / s, a[1] / s, a[2] / s]


def dot(a, b):
    return f32(f32(a[0] * b[0] + a[1] * b[1]) + a[2] * b[2])


class NpSim:
    def __init__(self, box):
        self.box = [f32(box)] * 3 if np.isscalar(box) else [f32(b) for b in box]
        self.res = [int(math.ceil(float(f32(b / H)))) for b in self.box]  # src/CBaseParticleSimulator.cpp:27-31
        self.cells = {}
        self.pos, self.vel, self.acc, self.rho, self.prs = [], [], [], [], []
        self.gravity = [f32(0), f32(-9.80665), f32(0)]
        h = float(H)
        self.poly6 = 315.0 / (64.0 * math.pi * math.pow(h, 9))  # src/CCPUParticleSimulator.cpp:11
        self.spiky = -45.0 / (math.pi * math.pow(h, 6))  # :19
        self.visc = 45.0 / (math.pi * math.pow(h, 6))  # :26
        self.h2 = float(f32(H * H))  # :12
        mn = [-(b / f32(2)) for b in self.box]
        mx = [b / f32(2) for b in self.box]
        z = f32(0)
        self.walls = [  # include/CCollisionGeometry.h:79-120
            ([f32(-1), z, z], [mn[0], z, z]), ([z, f32(-1), z], [z, mn[1], z]), ([z, z, f32(-1)], [z, z, mn[2]]),
            ([f32(1), z, z], [mx[0], z, z]), ([z, f32(1), z], [z, mx[1], z]), ([z, z, f32(1)], [z, z, mx[2]]),
        ]

        self.faces = []  # (normal, [v0, v1, v2]) — optional collision mesh, unused by the reference's own step

    def set_faces(self, faces12):
        self.faces = []
        for f in np.asarray(faces12, dtype=np.float32).reshape(-1, 12):
            self.faces.append(([f32(c) for c in f[0:3]], [[f32(c) for c in f[3 + 3 * v:6 + 3 * v]] for v in range(3)]))

    def cell(self, x, y, z):
        return self.cells.setdefault((x, y, z), [])

    def add(self, x, y, z):
        self.pos.append([f32(x), f32(y), f32(z)])
        self.vel.append([f32(0)] * 3)
        self.acc.append([f32(0)] * 3)
        self.rho.append(f32(0))
        self.prs.append(f32(0))
        self.cell(0, 0, 0).append(len(self.pos) - 1)  # src/CBaseParticleSimulator.cpp:69-72

    def setup_dam_break(self):  # src/CBaseParticleSimulator.cpp:38-58
        hp = float(f32(H / f32(2)))
        off = [-(b) / f32(2) for b in self.box]
        y = f32(0)
        while y < self.box[1]:
            x = f32(0)
            while float(x) < float(self.box[0]) / 4.0:
                z = f32(0)
                while z < self.box[2]:
                    self.add(x + off[0], y + off[1], z + off[2])
                    z = f32(float(z) + hp)
                x = f32(float(x) + hp)
            y = f32(float(y) + hp)
        return self

    def update_grid(self):  # src/CCPUParticleSimulator.cpp:32-91
        for x in range(self.res[0]):
            for y in range(self.res[1]):
                for z in range(self.res[2]):
                    plist = self.cell(x, y, z)
                    p = 0
                    while p < len(plist):
                        i = plist[p]
                        c = []
                        for a in range(3):
                            v = int(math.floor((float(self.pos[i][a]) + float(self.box[a]) / 2.0) / float(H)))
                            c.append(min(max(v, 0), self.res[a] - 1))
                        if (x, y, z) != tuple(c):
                            self.cell(*c).append(i)
                            plist[p] = plist[-1]
                            plist.pop()
                            continue  # redo this index
                        p += 1

    def neighbours_of(self, x, y, z):
        for ox in (-1, 0, 1):
            if x + ox < 0:
                continue
            if x + ox >= self.res[0]:
                break
            for oy in (-1, 0, 1):
                if y + oy < 0:
                    continue
                if y + oy >= self.res[1]:
                    break
                for oz in (-1, 0, 1):
                    if z + oz < 0:
                        continue
                    if z + oz >= self.res[2]:
                        break
                    yield from self.cells.get((x + ox, y + oy, z + oz), [])

    def each_particle(self):
        for x in range(self.res[0]):
            for y in range(self.res[1]):
                for z in range(self.res[2]):
                    for i in self.cells.get((x, y, z), []):
                        yield x, y, z, i

    def density_pressure(self):  # src/CCPUParticleSimulator.cpp:93-141
        self.nb = {}
        for x, y, z, i in self.each_particle():
            rho = f32(0)
            nb = []
            for j in self.neighbours_of(x, y, z):
                d = vsub(self.pos[i], self.pos[j])
                r2 = float(dot(d, d))
                if r2 <= float(f32(H * H)):
                    rho = f32(float(rho) + self.poly6 * math.pow(self.h2 - r2, 3))
                    nb.append(j)
            self.nb[i] = sorted(nb)
            rho = f32(rho * MASS)
            self.rho[i] = rho
            self.prs[i] = f32(GAS_STIFFNESS * f32(rho - REST_DENSITY))

    def wall_bounce(self, p, v):  # src/CCollisionGeometry.cpp:117-133
        acc = [f32(0)] * 3
        for normal, wpos in self.walls:
            inv = vmul(normal, -1.0)
            d = float(dot(vsub(wpos, p), inv)) + 0.01
            if d > 0.0:
                acc = vadd(acc, vmul(vmul(inv, WALL_K), d))
                acc = vadd(acc, vmul(inv, WALL_DAMPING * float(dot(v, inv))))
        return acc

    @staticmethod
    def qt_normalized(a):  # QVector3D::normalize() of Qt 5 (third-party; restated from its published source)
        ln = float(a[0]) * float(a[0]) + float(a[1]) * float(a[1]) + float(a[2]) * float(a[2])
        if abs(ln - 1.0) <= 1e-12 or abs(ln) <= 1e-12:
            return list(a)
        ln = math.sqrt(ln)
        return [f32(float(c) / ln) for c in a]

    def mesh_bounce(self, p, v):  # src/CCollisionGeometry.cpp:97-115
        acc = [f32(0)] * 3
        for normal, verts in self.faces:
            for q in verts:
                inv = self.qt_normalized(vmul(normal, -1.0))
                d = float(dot(vsub(q, p), inv)) + 0.01
                if d > 0.0:
                    acc = vadd(acc, vmul(vmul(inv, 5000.0), d))
                    acc = vadd(acc, vmul(inv, -0.9 * float(dot(v, inv))))
        return acc

    def forces(self):  # src/CCPUParticleSimulator.cpp:143-203
        for x, y, z, i in self.each_particle():
            fg = vmul(self.gravity, self.rho[i])
            fp, fv = [f32(0)] * 3, [f32(0)] * 3
            for j in self.neighbours_of(x, y, z):
                d = vsub(self.pos[i], self.pos[j])
                r2 = float(dot(d, d))
                if r2 <= float(f32(H * H)) and i != j:
                    r = math.sqrt(r2)
                    grad = vdiv(vmul(d, self.spiky * math.pow(float(H) - r, 2)), r)
                    lap = self.visc * (float(H) - r)
                    s = float(self.prs[i]) / math.pow(float(self.rho[i]), 2) + float(self.prs[j]) / math.pow(float(self.rho[j]), 2)
                    fp = vadd(fp, vmul(grad, s))
                    fv = vadd(fv, vdiv(vmul(vsub(self.vel[j], self.vel[i]), lap), self.rho[j]))
            fp = vmul(fp, f32(-MASS * self.rho[i]))
            fv = vmul(fv, f32(VISCOSITY * MASS))
            a = vdiv(vadd(vadd(fp, fv), fg), self.rho[i])
            a = vadd(a, self.wall_bounce(self.pos[i], self.vel[i]))
            if self.faces:  # extension point: inverseBounce added the way the bounding-box term is
                a = vadd(a, self.mesh_bounce(self.pos[i], self.vel[i]))
            self.acc[i] = a

    def integrate(self):  # src/CCPUParticleSimulator.cpp:211-229
        for i in range(len(self.pos)):
            new = vadd(vadd(self.pos[i], vmul(self.vel[i], DT)), vmul(vmul(self.acc[i], DT), DT))
            self.vel[i] = vdiv(vsub(new, self.pos[i]), DT)
            self.pos[i] = new

    def step(self):  # src/CBaseParticleSimulator.cpp:116-144
        self.update_grid()
        self.density_pressure()
        self.forces()
        self.integrate()

    def arrays(self):
        return (np.array(self.pos, dtype=np.float32), np.array(self.vel, dtype=np.float32), np.array(self.acc, dtype=np.float32),
                np.array(self.rho, dtype=np.float32), np.array(self.prs, dtype=np.float32))
