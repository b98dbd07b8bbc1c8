"""Oracle (TEST INFRASTRUCTURE, see oracle/__init__.py): CPU restatement of the
reference's Lie-group backend `lietorch_backends`.

The reference C++ needs Eigen 3.4.0 (README.md:71-73), absent here, so the
formulas below follow the reference headers line by line:

    SO3    devo/lietorch/include/so3.h:13-225
    RxSO3  devo/lietorch/include/rxso3.h:12-320
    SE3    devo/lietorch/include/se3.h:13-225
    Sim3   devo/lietorch/include/sim3.h:16-213
    ops    devo/lietorch/src/lietorch_gpu.cu:20-294  (forward/backward per op)
    EPS    devo/lietorch/include/common.h:7  (1e-6)

Eigen semantics relied on: quaternion coefficient order (x,y,z,w), Hamilton
product, `normalize()` on every construction from data or from a quaternion
(so3.h:31-37, rxso3.h:30-37), `toRotationMatrix()`, 3x3 `inverse()`
(sim3.h:151).  Known reference quirks are reproduced on purpose:
  * Sim3::left_jacobian drops its 1/720 term (stray ';', sim3.h:177-178).

All functions take/return 2-D torch CPU tensors [batch, dim] in fp32 or fp64.
The module-level functions at the bottom mirror the 19 pybind entry points of
devo/lietorch/src/lietorch.cpp:286-316, so this module can be installed as
`lietorch_backends` under the reference's own Python wrappers.
"""
import math

import torch

EPS = 1e-6
PI = 3.14159265358979323846


# ----------------------------------------------------------------------------- helpers
def _col(x):
    return x.unsqueeze(-1)


def hat(v):
    """so3.h:88-96"""
    o = torch.zeros_like(v[:, 0])
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=-1).view(-1, 3, 3)


def eye(n, like, b=None):
    I = torch.eye(n, dtype=like.dtype, device=like.device)
    if b is None:
        b = like.shape[0]
    return I.unsqueeze(0).repeat(b, 1, 1)


def qnormalize(q):
    return q / torch.sqrt((q * q).sum(-1, keepdim=True))


def qmul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz], dim=-1)


def qconj(q):
    return torch.cat([-q[:, :3], q[:, 3:]], dim=-1)


def qrot(q, p):
    """so3.h:52-57"""
    qv, w = q[:, :3], q[:, 3:]
    uv = torch.linalg.cross(qv, p)
    uv = uv + uv
    return p + w * uv + torch.linalg.cross(qv, uv)


def qmat(q):
    """Eigen::Quaternion::toRotationMatrix"""
    x, y, z, w = q.unbind(-1)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return torch.stack([
        1 - (tyy + tzz), txy - twz, txz + twy,
        txy + twz, 1 - (txx + tzz), tyz - twx,
        txz - twy, tyz + twx, 1 - (txx + tyy)], dim=-1).view(-1, 3, 3)


def _mv(M, v):
    return torch.matmul(M, v.unsqueeze(-1)).squeeze(-1)


def _vm(v, M):
    return torch.matmul(v.unsqueeze(-2), M).squeeze(-2)


def _exp_quat(phi):
    """so3.h:141-157 / rxso3.h:177-196 (before the normalising constructor)"""
    theta2 = (phi * phi).sum(-1)
    theta = torch.sqrt(theta2)
    small = theta < EPS
    theta4 = theta2 * theta2
    ts = torch.where(small, torch.ones_like(theta), theta)
    imag = torch.where(small, 0.5 - (1.0 / 48.0) * theta2 + (1.0 / 3840.0) * theta4,
                       torch.sin(0.5 * ts) / ts)
    real = torch.where(small, 1.0 - (1.0 / 8.0) * theta2 + (1.0 / 384.0) * theta4,
                       torch.cos(0.5 * ts))
    q = torch.cat([_col(imag) * phi, _col(real)], dim=-1)
    return qnormalize(q)


def _log_quat(q):
    """so3.h:106-139 / rxso3.h:132-166"""
    qv, w = q[:, :3], q[:, 3]
    sq = (qv * qv).sum(-1)
    small = sq < EPS * EPS
    n = torch.sqrt(torch.where(small, torch.ones_like(sq), sq))
    wz = w.abs() < EPS
    ws = torch.where(wz, torch.ones_like(w), w)
    f_small = 2.0 / ws - (2.0 / 3.0) * sq / (ws * ws * ws)
    f_wz = torch.where(w > 0, PI / n, -PI / n)
    f_gen = 2.0 * torch.atan(n / ws) / n
    f = torch.where(small, f_small, torch.where(wz, f_wz, f_gen))
    return _col(f) * qv


def _so3_left_jacobian(phi):
    """so3.h:159-177"""
    Phi = hat(phi)
    Phi2 = torch.matmul(Phi, Phi)
    theta2 = (phi * phi).sum(-1)
    theta = torch.sqrt(theta2)
    small = theta < EPS
    t2 = torch.where(small, torch.ones_like(theta2), theta2)
    t = torch.where(small, torch.ones_like(theta), theta)
    c1 = torch.where(small, 0.5 - (1.0 / 24.0) * theta2, (1.0 - torch.cos(t)) / t2)
    c2 = torch.where(small, 1.0 / 6.0 - (1.0 / 120.0) * theta2, (t - torch.sin(t)) / (t2 * t))
    return eye(3, phi) + c1.view(-1, 1, 1) * Phi + c2.view(-1, 1, 1) * Phi2


def _so3_left_jacobian_inverse(phi):
    """so3.h:179-195"""
    Phi = hat(phi)
    Phi2 = torch.matmul(Phi, Phi)
    theta2 = (phi * phi).sum(-1)
    theta = torch.sqrt(theta2)
    small = theta < EPS
    t = torch.where(small, torch.ones_like(theta), theta)
    h = 0.5 * t
    c2 = torch.where(small, torch.full_like(theta, 1.0 / 12.0),
                     (1.0 - t * torch.cos(h) / (2.0 * torch.sin(h))) / (t * t))
    return eye(3, phi) - 0.5 * Phi + c2.view(-1, 1, 1) * Phi2


def _blk(rows):
    """assemble a block matrix from a list of lists of [b,r,c] tensors"""
    return torch.cat([torch.cat(r, dim=-1) for r in rows], dim=-2)


# ----------------------------------------------------------------------------- groups
class SO3:
    gid, K, N = 1, 3, 4

    @staticmethod
    def from_data(d):
        return qnormalize(d)

    @staticmethod
    def Exp(a):
        return _exp_quat(a)

    @staticmethod
    def Log(X):
        return _log_quat(X)

    @staticmethod
    def inv(X):
        return qnormalize(qconj(X))

    @staticmethod
    def mul(X, Y):
        return qnormalize(qmul(X, Y))

    @staticmethod
    def act(X, p):
        return qrot(X, p)

    @staticmethod
    def act4(X, p):
        return torch.cat([qrot(X, p[:, :3]), p[:, 3:]], dim=-1)

    @staticmethod
    def Adj(X):
        return qmat(X)

    @staticmethod
    def adj(a):
        return hat(a)

    @staticmethod
    def Matrix4(X):
        T = eye(4, X)
        T[:, :3, :3] = qmat(X)
        return T

    left_jacobian = staticmethod(_so3_left_jacobian)
    left_jacobian_inverse = staticmethod(_so3_left_jacobian_inverse)

    @staticmethod
    def projector(X):
        """so3.h:73-82"""
        qv, w = X[:, :3], X[:, 3]
        J = torch.zeros(X.shape[0], 4, 4, dtype=X.dtype)
        J[:, :3, :3] = 0.5 * (w.view(-1, 1, 1) * eye(3, X) + hat(-qv))
        J[:, 3, :3] = 0.5 * (-qv)
        return J

    @staticmethod
    def act_jacobian(p):
        return hat(-p)

    @staticmethod
    def act4_jacobian(p):
        J = torch.zeros(p.shape[0], 4, 3, dtype=p.dtype)
        J[:, :3, :3] = hat(-p[:, :3])
        return J


class RxSO3:
    gid, K, N = 2, 4, 5

    @staticmethod
    def from_data(d):
        return torch.cat([qnormalize(d[:, :4]), d[:, 4:]], dim=-1)

    @staticmethod
    def Exp(a):
        return torch.cat([_exp_quat(a[:, :3]), torch.exp(a[:, 3:])], dim=-1)

    @staticmethod
    def Log(X):
        return torch.cat([_log_quat(X[:, :4]), torch.log(X[:, 4:])], dim=-1)

    @staticmethod
    def inv(X):
        return torch.cat([qnormalize(qconj(X[:, :4])), 1.0 / X[:, 4:]], dim=-1)

    @staticmethod
    def mul(X, Y):
        return torch.cat([qnormalize(qmul(X[:, :4], Y[:, :4])), X[:, 4:] * Y[:, 4:]], dim=-1)

    @staticmethod
    def act(X, p):
        return X[:, 4:] * qrot(X[:, :4], p)

    @staticmethod
    def act4(X, p):
        return torch.cat([RxSO3.act(X, p[:, :3]), p[:, 3:]], dim=-1)

    @staticmethod
    def Rotation(X):
        return qmat(X[:, :4])

    @staticmethod
    def Matrix(X):
        return X[:, 4].view(-1, 1, 1) * qmat(X[:, :4])

    @staticmethod
    def Adj(X):
        A = eye(4, X)
        A[:, :3, :3] = qmat(X[:, :4])
        return A

    @staticmethod
    def adj(a):
        A = torch.zeros(a.shape[0], 4, 4, dtype=a.dtype)
        A[:, :3, :3] = hat(a[:, :3])
        return A

    @staticmethod
    def Matrix4(X):
        T = eye(4, X)
        T[:, :3, :3] = RxSO3.Matrix(X)
        return T

    @staticmethod
    def left_jacobian(a):
        J = eye(4, a)
        J[:, :3, :3] = _so3_left_jacobian(a[:, :3])
        return J

    @staticmethod
    def left_jacobian_inverse(a):
        J = eye(4, a)
        J[:, :3, :3] = _so3_left_jacobian_inverse(a[:, :3])
        return J

    @staticmethod
    def projector(X):
        """rxso3.h:84-100"""
        qv, w, s = X[:, :3], X[:, 3], X[:, 4]
        J = torch.zeros(X.shape[0], 5, 5, dtype=X.dtype)
        J[:, :3, :3] = 0.5 * (w.view(-1, 1, 1) * eye(3, X) + hat(-qv))
        J[:, 3, :3] = 0.5 * (-qv)
        J[:, 4, 3] = s
        return J

    @staticmethod
    def act_jacobian(p):
        return torch.cat([hat(-p), p.unsqueeze(-1)], dim=-1)

    @staticmethod
    def act4_jacobian(p):
        J = torch.zeros(p.shape[0], 4, 4, dtype=p.dtype)
        J[:, :3, :3] = hat(-p[:, :3])
        J[:, :3, 3] = p[:, :3]
        return J

    @staticmethod
    def calcW(a):
        """rxso3.h:198-241"""
        phi, sigma = a[:, :3], a[:, 3]
        theta = torch.sqrt((phi * phi).sum(-1))
        Phi = hat(phi)
        Phi2 = torch.matmul(Phi, Phi)
        scale = torch.exp(sigma)
        one = torch.ones_like(sigma)
        ssm = sigma.abs() < EPS
        tsm = theta.abs() < EPS
        sg = torch.where(ssm, one, sigma)
        th = torch.where(tsm, one, theta)
        th2 = th * th
        # sigma small
        A00 = 0.5 * one
        B00 = one / 6.0
        A01 = (one - torch.cos(th)) / th2
        B01 = (th - torch.sin(th)) / (th2 * th)
        # sigma not small
        C1 = (scale - one) / sg
        sg2 = sg * sg
        A10 = ((sg - one) * scale + one) / sg2
        B10 = (scale * 0.5 * sg2 + scale - one - sg * scale) / (sg2 * sg)
        a_ = scale * torch.sin(th)
        b_ = scale * torch.cos(th)
        c_ = th2 + sg * sg
        A11 = (a_ * sg + (one - b_) * th) / (th * c_)
        B11 = (C1 - ((b_ - one) * sg + a_ * th) / c_) * one / th2
        A = torch.where(ssm, torch.where(tsm, A00, A01), torch.where(tsm, A10, A11))
        B = torch.where(ssm, torch.where(tsm, B00, B01), torch.where(tsm, B10, B11))
        C = torch.where(ssm, one, C1)
        return A.view(-1, 1, 1) * Phi + B.view(-1, 1, 1) * Phi2 + C.view(-1, 1, 1) * eye(3, a)


class SE3:
    gid, K, N = 3, 6, 7

    @staticmethod
    def from_data(d):
        return torch.cat([d[:, :3], qnormalize(d[:, 3:7])], dim=-1)

    @staticmethod
    def Exp(a):
        tau, phi = a[:, :3], a[:, 3:6]
        q = _exp_quat(phi)
        t = _mv(_so3_left_jacobian(phi), tau)
        return torch.cat([t, q], dim=-1)

    @staticmethod
    def Log(X):
        phi = _log_quat(X[:, 3:7])
        tau = _mv(_so3_left_jacobian_inverse(phi), X[:, :3])
        return torch.cat([tau, phi], dim=-1)

    @staticmethod
    def inv(X):
        qi = SO3.inv(X[:, 3:7])
        return torch.cat([-qrot(qi, X[:, :3]), qi], dim=-1)

    @staticmethod
    def mul(X, Y):
        q = SO3.mul(X[:, 3:7], Y[:, 3:7])
        t = X[:, :3] + qrot(X[:, 3:7], Y[:, :3])
        return torch.cat([t, q], dim=-1)

    @staticmethod
    def act(X, p):
        return qrot(X[:, 3:7], p) + X[:, :3]

    @staticmethod
    def act4(X, p):
        return torch.cat([qrot(X[:, 3:7], p[:, :3]) + X[:, :3] * p[:, 3:], p[:, 3:]], dim=-1)

    @staticmethod
    def Adj(X):
        R = qmat(X[:, 3:7])
        tx = hat(X[:, :3])
        Z = torch.zeros_like(R)
        return _blk([[R, torch.matmul(tx, R)], [Z, R]])

    @staticmethod
    def adj(a):
        Tau, Phi = hat(a[:, :3]), hat(a[:, 3:6])
        Z = torch.zeros_like(Phi)
        return _blk([[Phi, Tau], [Z, Phi]])

    @staticmethod
    def Matrix4(X):
        T = eye(4, X)
        T[:, :3, :3] = qmat(X[:, 3:7])
        T[:, :3, 3] = X[:, :3]
        return T

    @staticmethod
    def calcQ(a):
        """se3.h:142-172"""
        Tau, Phi = hat(a[:, :3]), hat(a[:, 3:6])
        phi = a[:, 3:6]
        theta = torch.sqrt((phi * phi).sum(-1))
        t2 = theta * theta
        t4 = t2 * t2
        small = theta < EPS
        th = torch.where(small, torch.ones_like(theta), theta)
        h2 = th * th
        h4 = h2 * h2
        c1 = torch.where(small, 1.0 / 6.0 - (1.0 / 120.0) * t2, (th - torch.sin(th)) / (h2 * th))
        c2 = torch.where(small, 1.0 / 24.0 - (1.0 / 720.0) * t2, (h2 + 2 * torch.cos(th) - 2) / (2 * h4))
        c3 = torch.where(small, 1.0 / 120.0 - (1.0 / 2520.0) * t2,
                         (2 * th - 3 * torch.sin(th) + th * torch.cos(th)) / (2 * h4 * th))
        mm = torch.matmul
        PT, TP = mm(Phi, Tau), mm(Tau, Phi)
        PTP = mm(PT, Phi)
        PP = mm(Phi, Phi)
        Q = 0.5 * Tau + c1.view(-1, 1, 1) * (PT + TP + PTP) \
            + c2.view(-1, 1, 1) * (mm(PP, Tau) + mm(Tau, PP) - 3 * PTP) \
            + c3.view(-1, 1, 1) * (mm(PTP, Phi) + mm(Phi, PTP))
        return Q

    @staticmethod
    def left_jacobian(a):
        J = _so3_left_jacobian(a[:, 3:6])
        Q = SE3.calcQ(a)
        Z = torch.zeros_like(J)
        return _blk([[J, Q], [Z, J]])

    @staticmethod
    def left_jacobian_inverse(a):
        Ji = _so3_left_jacobian_inverse(a[:, 3:6])
        Q = SE3.calcQ(a)
        Z = torch.zeros_like(Ji)
        return _blk([[Ji, -torch.matmul(torch.matmul(Ji, Q), Ji)], [Z, Ji]])

    @staticmethod
    def projector(X):
        """se3.h:116-124"""
        J = torch.zeros(X.shape[0], 7, 7, dtype=X.dtype)
        J[:, :3, :3] = eye(3, X)
        J[:, :3, 3:6] = hat(-X[:, :3])
        J[:, 3:7, 3:7] = SO3.projector(X[:, 3:7])
        return J

    @staticmethod
    def act_jacobian(p):
        return torch.cat([eye(3, p), hat(-p)], dim=-1)

    @staticmethod
    def act4_jacobian(p):
        J = torch.zeros(p.shape[0], 4, 6, dtype=p.dtype)
        J[:, :3, :3] = p[:, 3].view(-1, 1, 1) * eye(3, p)
        J[:, :3, 3:6] = hat(-p[:, :3])
        return J


class Sim3:
    gid, K, N = 4, 7, 8

    @staticmethod
    def from_data(d):
        return torch.cat([d[:, :3], qnormalize(d[:, 3:7]), d[:, 7:]], dim=-1)

    @staticmethod
    def Exp(a):
        tau, ps = a[:, :3], a[:, 3:7]
        r = RxSO3.Exp(ps)
        t = _mv(RxSO3.calcW(ps), tau)
        return torch.cat([t, r], dim=-1)

    @staticmethod
    def Log(X):
        ps = RxSO3.Log(X[:, 3:8])
        W = RxSO3.calcW(ps)
        tau = _mv(torch.linalg.inv(W), X[:, :3])
        return torch.cat([tau, ps], dim=-1)

    @staticmethod
    def inv(X):
        ri = RxSO3.inv(X[:, 3:8])
        return torch.cat([-RxSO3.act(ri, X[:, :3]), ri], dim=-1)

    @staticmethod
    def mul(X, Y):
        r = RxSO3.mul(X[:, 3:8], Y[:, 3:8])
        t = X[:, :3] + RxSO3.act(X[:, 3:8], Y[:, :3])
        return torch.cat([t, r], dim=-1)

    @staticmethod
    def act(X, p):
        return RxSO3.act(X[:, 3:8], p) + X[:, :3]

    @staticmethod
    def act4(X, p):
        return torch.cat([RxSO3.act(X[:, 3:8], p[:, :3]) + p[:, 3:] * X[:, :3], p[:, 3:]], dim=-1)

    @staticmethod
    def Adj(X):
        """sim3.h:85-98"""
        t = X[:, :3]
        sR = RxSO3.Matrix(X[:, 3:8])
        R = RxSO3.Rotation(X[:, 3:8])
        A = eye(7, X)
        A[:, :3, :3] = sR
        A[:, :3, 3:6] = torch.matmul(hat(t), R)
        A[:, :3, 6] = -t
        A[:, 3:6, 3:6] = R
        return A

    @staticmethod
    def adj(a):
        """sim3.h:124-141"""
        tau, phi, sigma = a[:, :3], a[:, 3:6], a[:, 6]
        A = torch.zeros(a.shape[0], 7, 7, dtype=a.dtype)
        A[:, :3, :3] = hat(phi) + sigma.view(-1, 1, 1) * eye(3, a)
        A[:, :3, 3:6] = hat(tau)
        A[:, :3, 6] = -tau
        A[:, 3:6, 3:6] = hat(phi)
        return A

    @staticmethod
    def Matrix4(X):
        T = eye(4, X)
        T[:, :3, :3] = RxSO3.Matrix(X[:, 3:8])
        T[:, :3, 3] = X[:, :3]
        return T

    @staticmethod
    def left_jacobian(a):
        """sim3.h:167-179 -- the 1/720 term is dropped by a stray ';' in the reference"""
        Xi = Sim3.adj(a)
        Xi2 = torch.matmul(Xi, Xi)
        Xi4 = torch.matmul(Xi2, Xi2)
        return eye(7, a) + 0.5 * Xi + (1.0 / 6.0) * Xi2 + (1.0 / 24.0) * torch.matmul(Xi, Xi2) + (1.0 / 120.0) * Xi4

    @staticmethod
    def left_jacobian_inverse(a):
        """sim3.h:181-191"""
        Xi = Sim3.adj(a)
        Xi2 = torch.matmul(Xi, Xi)
        Xi4 = torch.matmul(Xi2, Xi2)
        return eye(7, a) - 0.5 * Xi + (1.0 / 12.0) * Xi2 - (1.0 / 720.0) * Xi4

    @staticmethod
    def projector(X):
        """sim3.h:74-83"""
        t = X[:, :3]
        J = torch.zeros(X.shape[0], 8, 8, dtype=X.dtype)
        J[:, :3, :3] = eye(3, X)
        J[:, :3, 3:6] = hat(-t)
        J[:, :3, 6] = t
        J[:, 3:8, 3:8] = RxSO3.projector(X[:, 3:8])
        return J

    @staticmethod
    def act_jacobian(p):
        return torch.cat([eye(3, p), hat(-p), p.unsqueeze(-1)], dim=-1)

    @staticmethod
    def act4_jacobian(p):
        J = torch.zeros(p.shape[0], 4, 7, dtype=p.dtype)
        J[:, :3, :3] = p[:, 3].view(-1, 1, 1) * eye(3, p)
        J[:, :3, 3:6] = hat(-p[:, :3])
        J[:, :3, 6] = p[:, :3]
        return J


GROUPS = {1: SO3, 2: RxSO3, 3: SE3, 4: Sim3}


def _pad(g, v, like):
    """gradient w.r.t. a group element: K-vector in the first K of N slots
    (lietorch_gpu.cu:41-42,120-123; outputs are zero-initialised :315,345,...)"""
    out = torch.zeros(like.shape[0], g.N, dtype=v.dtype)
    out[:, :g.K] = v
    return out


# ----------------------------------------------------------------------------- the 19 entry points
# (devo/lietorch/src/lietorch.cpp:286-316; kernels lietorch_gpu.cu:20-294)
def expm(gid, a):
    return GROUPS[gid].Exp(a)


def expm_backward(gid, grad, a):
    g = GROUPS[gid]
    return [_vm(grad[:, :g.K], g.left_jacobian(a))]


def logm(gid, X):
    g = GROUPS[gid]
    return g.Log(g.from_data(X))


def logm_backward(gid, grad, X):
    g = GROUPS[gid]
    a = g.Log(g.from_data(X))
    return [_pad(g, _vm(grad, g.left_jacobian_inverse(a)), X)]


def inv(gid, X):
    g = GROUPS[gid]
    return g.inv(g.from_data(X))


def inv_backward(gid, grad, X):
    g = GROUPS[gid]
    Y = g.inv(g.from_data(X))
    return [_pad(g, -_vm(grad[:, :g.K], g.Adj(Y)), X)]


def mul(gid, X, Y):
    g = GROUPS[gid]
    return g.mul(g.from_data(X), g.from_data(Y))


def mul_backward(gid, grad, X, Y):
    g = GROUPS[gid]
    dZ = grad[:, :g.K]
    return [_pad(g, dZ, X), _pad(g, _vm(dZ, g.Adj(g.from_data(X))), Y)]


def adj(gid, X, a):
    g = GROUPS[gid]
    return _mv(g.Adj(g.from_data(X)), a)


def adj_backward(gid, grad, X, a):
    g = GROUPS[gid]
    A = g.Adj(g.from_data(X))
    b = _mv(A, a)
    return [_pad(g, -_vm(grad, g.adj(b)), X), _vm(grad, A)]


def adjT(gid, X, a):
    g = GROUPS[gid]
    return _mv(g.Adj(g.from_data(X)).transpose(-1, -2), a)


def adjT_backward(gid, grad, X, a):
    g = GROUPS[gid]
    A = g.Adj(g.from_data(X))
    Adb = _mv(A, grad)
    return [_pad(g, -_vm(a, g.adj(Adb)), X), Adb]


def act(gid, X, p):
    g = GROUPS[gid]
    return g.act(g.from_data(X), p)


def act_backward(gid, grad, X, p):
    g = GROUPS[gid]
    Xn = g.from_data(X)
    q = g.act(Xn, p)
    dp = _vm(grad, g.Matrix4(Xn)[:, :3, :3])
    return [_pad(g, _vm(grad, g.act_jacobian(q)), X), dp]


def act4(gid, X, p):
    g = GROUPS[gid]
    return g.act4(g.from_data(X), p)


def act4_backward(gid, grad, X, p):
    g = GROUPS[gid]
    Xn = g.from_data(X)
    q = g.act4(Xn, p)
    dp = _vm(grad, g.Matrix4(Xn))
    return [_pad(g, _vm(grad, g.act4_jacobian(q)), X), dp]


def as_matrix(gid, X):
    g = GROUPS[gid]
    return g.Matrix4(g.from_data(X))


def projector(gid, X):
    g = GROUPS[gid]
    return g.projector(g.from_data(X))


def Jinv(gid, X, a):
    g = GROUPS[gid]
    return _mv(g.left_jacobian_inverse(g.Log(g.from_data(X))), a)


ENTRY_POINTS = ["expm", "expm_backward", "logm", "logm_backward", "inv", "inv_backward",
                "mul", "mul_backward", "adj", "adj_backward", "adjT", "adjT_backward",
                "act", "act_backward", "act4", "act4_backward", "as_matrix", "projector", "Jinv"]
assert len(ENTRY_POINTS) == 19 and math.isclose(PI, math.pi)
