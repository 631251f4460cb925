"""TEST INFRASTRUCTURE — an independent model of what `registerFeatures` hands to Ceres 2.2.0.

The reference builds one `ceres::Problem` per outer iteration (registration-inl.h:30-56): parameter blocks
`rotation.coeffs()` (4, `QuaternionManifold`) and `translation` (3, `EuclideanManifold<3>`), one
`AutoDiffCostFunction<EdgeCostFunction,1,4,3>` / `<PlaneCostFunction,1,4,3>` per association with `HuberLoss(1.0)`,
solved with TRUST_REGION / LEVENBERG_MARQUARDT / DENSE_QR, `max_num_iterations = 4`.

This file restates that a SECOND time, as differently from oracle/loam_oracle.c as possible, so that the two can be
checked against each other (tests/test_oracle_jacobians.py):

  * residuals are the LITERAL functor expressions (registration-inl.h:92-117, geometry-inl.h:21-33, Eigen's
    `Quaternion * Vector3`: `v + w*uv + u x uv` with `uv = 2 (u x v)`) written in torch float64, and every derivative
    comes from torch autograd — the stand-in for `ceres::Jet` autodiff; the oracle uses hand-derived analytic forms;
  * the linear solve is a dense QR of the stacked `[J ; diag(D)]` matrix through LAPACK (numpy.linalg.qr), as Ceres'
    DENSE_QR does; the oracle carries its own Householder loop and the kernel the 6x6 damped normal equations;
  * the minimizer follows TrustRegionMinimizer / LevenbergMarquardtStrategy / TrustRegionStepEvaluator method by
    method (names kept) instead of the oracle's single flattened loop.

It is NOT Ceres: both restatements come from the published algorithm (Ceres is absent from /root/reference and from
this image), so agreement pins derivations and arithmetic, not the behaviour of the real library — see DESIGN.md §7.
"""
from __future__ import annotations

import numpy as np
import torch

DBL_MIN = np.finfo(np.float64).tiny
DBL_MAX = np.finfo(np.float64).max


# ------------------------------------------------------------------------------------------------ functors
def eigen_rotate(q_xyzw: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Eigen::QuaternionBase::_transformVector (no normalisation): uv = 2 (u x v); v + w uv + u x uv."""
    u, w = q_xyzw[:3], q_xyzw[3]
    uv = torch.linalg.cross(u.expand_as(v), v, dim=-1)
    uv = uv + uv
    return v + w * uv + torch.linalg.cross(u.expand_as(v), uv, dim=-1)


def residuals(x: torch.Tensor, is_plane: torch.Tensor, P: torch.Tensor, A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """All residuals at the 7 ambient parameters x = (qx qy qz qw | tx ty tz) (Eigen coeffs() memory order).

    EdgeCostFunction::operator()  (registration-inl.h:92-103): pointToLineDistance(q * p + t, a, b)
    PlaneCostFunction::operator() (registration-inl.h:106-117): pointToPlaneDistance(q * p + t, n, d)
    """
    pt = eigen_rotate(x[:4], P) + x[4:]
    # geometry-inl.h:21-27 : ((p - a).cross(p - b)).norm() / (a - b).norm()
    num = torch.linalg.cross(pt - A, pt - B, dim=-1).norm(dim=-1)
    den = (A - B).norm(dim=-1)
    edge = num / torch.where(is_plane, torch.ones_like(den), den)
    # geometry-inl.h:30-33 : abs(n.dot(p) - d)
    plane = ((A * pt).sum(-1) - B[:, 0]).abs()
    return torch.where(is_plane, plane, edge)


def ambient_jacobian(x, is_plane, P, A, B):
    """(r [M], J [M,7]) with J from autograd — what AutoDiffCostFunction<., 1, 4, 3> produces per block."""
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    J = torch.autograd.functional.jacobian(lambda z: residuals(z, is_plane, P, A, B), xt, vectorize=True)
    r = residuals(xt, is_plane, P, A, B).detach()
    return r.numpy(), J.detach().numpy()


# ------------------------------------------------------------------------------------------------ manifold
def quaternion_manifold_plus(x4, delta3):
    """ceres::QuaternionManifold::Plus — the w-FIRST manifold applied to whatever 4 doubles it is given (the reference
    hands it Eigen's x,y,z,w memory: slot 0 plays "w")."""
    nd = np.sqrt(delta3[0] ** 2 + delta3[1] ** 2 + delta3[2] ** 2)
    if nd == 0.0:
        return np.array(x4, dtype=np.float64)
    s = np.sin(nd) / nd
    qd = np.array([np.cos(nd), s * delta3[0], s * delta3[1], s * delta3[2]])
    W, X, Y, Z = 0, 1, 2, 3
    o = np.empty(4)
    o[W] = qd[W] * x4[W] - qd[X] * x4[X] - qd[Y] * x4[Y] - qd[Z] * x4[Z]
    o[X] = qd[W] * x4[X] + qd[X] * x4[W] + qd[Y] * x4[Z] - qd[Z] * x4[Y]
    o[Y] = qd[W] * x4[Y] - qd[X] * x4[Z] + qd[Y] * x4[W] + qd[Z] * x4[X]
    o[Z] = qd[W] * x4[Z] + qd[X] * x4[Y] - qd[Y] * x4[X] + qd[Z] * x4[W]
    return o


def quaternion_manifold_plus_jacobian(x4):
    """ceres::QuaternionManifold::PlusJacobian, row = memory slot (CeresQuaternionOrder: kW = 0 .. kZ = 3)."""
    W, X, Y, Z = 0, 1, 2, 3
    j = np.zeros((4, 3))
    j[W] = [-x4[X], -x4[Y], -x4[Z]]
    j[X] = [x4[W], x4[Z], -x4[Y]]
    j[Y] = [-x4[Z], x4[W], x4[X]]
    j[Z] = [x4[Y], -x4[X], x4[W]]
    return j


def plus(x, delta6):
    """ProblemImpl Plus over both blocks: QuaternionManifold on slots 0-3, EuclideanManifold<3> on 4-6."""
    return np.concatenate([quaternion_manifold_plus(x[:4], delta6[:3]), x[4:] + delta6[3:]])


# ------------------------------------------------------------------------------------------------ loss + evaluator
def huber_rho(s, a=1.0):
    """ceres::HuberLoss::Evaluate → (rho, rho', rho'')."""
    b = a * a
    if s > b:
        r = np.sqrt(s)
        rho1 = max(DBL_MIN, a / r)
        return 2.0 * a * r - b, rho1, -rho1 / (2.0 * s)
    return s, 1.0, 0.0


class Problem:
    def __init__(self, is_plane, P, A, B):
        self.is_plane = torch.as_tensor(np.asarray(is_plane) != 0)
        self.P, self.A, self.B = (torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64)) for v in (P, A, B))
        self.M = len(self.P)

    def evaluate(self, x, want_jacobian):
        """ProgramEvaluator::Evaluate: cost = 1/2 sum rho(r^2); residuals and the tangent Jacobian corrected by
        ceres::Corrector (rho'' <= 0 for Huber: both scaled by sqrt(rho'))."""
        if want_jacobian:
            r, Ja = ambient_jacobian(x, self.is_plane, self.P, self.A, self.B)
            J = np.hstack([Ja[:, :4] @ quaternion_manifold_plus_jacobian(x[:4]), Ja[:, 4:]])
        else:
            with torch.no_grad():
                r = residuals(torch.tensor(x, dtype=torch.float64), self.is_plane, self.P, self.A, self.B).numpy()
            J = None
        cost = 0.0
        rc = r.copy()
        for i in range(self.M):
            rho, rho1, rho2 = huber_rho(r[i] * r[i])
            cost += 0.5 * rho
            assert rho2 <= 0.0
            sr = np.sqrt(rho1)
            rc[i] = sr * r[i]
            if J is not None:
                J[i] *= sr
        g = J.T @ rc if J is not None else None
        return cost, rc, J, g


# ------------------------------------------------------------------------------------------------ minimizer
class LevenbergMarquardtStrategy:
    def __init__(self):
        self.radius, self.max_radius, self.min_diagonal, self.max_diagonal = 1e4, 1e16, 1e-6, 1e32
        self.decrease_factor, self.reuse_diagonal, self.diagonal = 2.0, False, None

    def compute_step(self, J, r):
        if not self.reuse_diagonal:
            self.diagonal = np.clip((J * J).sum(0), self.min_diagonal, self.max_diagonal)
        D = np.sqrt(self.diagonal / self.radius)
        # DenseQRSolver: min | [J ; diag(D)] y - [r ; 0] |, then step = -y
        Aug = np.vstack([J, np.diag(D)])
        rhs = np.concatenate([r, np.zeros(J.shape[1])])
        Q, R = np.linalg.qr(Aug)
        y = np.linalg.solve(R, Q.T @ rhs)
        self.reuse_diagonal = True
        return -y if np.all(np.isfinite(y)) else None

    def step_accepted(self, q):
        self.radius = min(self.max_radius, self.radius / max(1.0 / 3.0, 1.0 - (2.0 * q - 1.0) ** 3))
        self.decrease_factor, self.reuse_diagonal = 2.0, False

    def step_rejected(self, q):
        self.radius /= self.decrease_factor
        self.decrease_factor *= 2.0
        self.reuse_diagonal = True

    def step_is_invalid(self):
        self.radius *= 0.5
        self.reuse_diagonal = True


def solve(problem: Problem, x0=None, max_num_iterations=4, arm_after_success=True, trace=None):
    """ceres::Solve(options{DENSE_QR, max_num_iterations = 4}, ...): TrustRegionMinimizer::Minimize, monotonic steps,
    jacobi_scaling, all tolerances at their 2.2.0 defaults.  Returns (x, iterations recorded, (cost0, cost))."""
    function_tolerance, gradient_tolerance, parameter_tolerance = 1e-6, 1e-10, 1e-8
    min_relative_decrease, min_trust_region_radius = 1e-3, 1e-32
    x = np.array([0, 0, 0, 1, 0, 0, 0.0]) if x0 is None else np.array(x0, dtype=np.float64)
    strategy = LevenbergMarquardtStrategy()

    # ---- IterationZero / EvaluateGradientAndJacobian
    x_norm = np.linalg.norm(x)
    x_cost, res, J, grad = problem.evaluate(x, True)
    cost0 = x_cost
    jacobian_scaling = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    J = J * jacobian_scaling
    gradient_max_norm = np.abs(x - plus(x, -grad)).max()
    iteration, step_is_successful, atleast_one_successful_step = 0, True, False
    num_consecutive_invalid_steps = 0
    while True:
        # ---- FinalizeIterationAndCheckIfMinimizerCanContinue
        if trace is not None:
            trace.append(dict(iteration=iteration, x=x.copy(), cost=x_cost, radius=strategy.radius,
                              successful=step_is_successful))
        if iteration >= max_num_iterations:
            break
        if step_is_successful and gradient_max_norm <= gradient_tolerance:
            break
        if strategy.radius <= min_trust_region_radius:
            break
        iteration += 1
        # ---- ComputeTrustRegionStep
        step = strategy.compute_step(J, res)
        step_is_valid = False
        if step is not None:
            model_residuals = J @ step
            model_cost_change = -model_residuals @ (res + model_residuals / 2.0)
            step_is_valid = model_cost_change > 0.0
        if not step_is_valid:  # HandleInvalidStep
            num_consecutive_invalid_steps += 1
            if num_consecutive_invalid_steps >= 5:
                break
            strategy.step_is_invalid()
            step_is_successful = False
            continue
        num_consecutive_invalid_steps = 0
        delta = step * jacobian_scaling
        # ---- ComputeCandidatePointAndEvaluateCost
        candidate_x = plus(x, delta)
        candidate_cost = problem.evaluate(candidate_x, False)[0]
        # ---- ParameterToleranceReached / FunctionToleranceReached (only once a step has succeeded, 2.2.0)
        if atleast_one_successful_step or not arm_after_success:
            if np.linalg.norm(x - candidate_x) <= parameter_tolerance * (x_norm + parameter_tolerance):
                break
            if abs(x_cost - candidate_cost) <= function_tolerance * x_cost:
                break
        # ---- IsStepSuccessful (TrustRegionStepEvaluator, monotonic: reference cost == current cost)
        relative_decrease = (x_cost - candidate_cost) / model_cost_change
        if relative_decrease > min_relative_decrease:  # HandleSuccessfulStep
            atleast_one_successful_step = True
            x = candidate_x
            x_norm = np.linalg.norm(x)
            x_cost, res, J, grad = problem.evaluate(x, True)
            J = J * jacobian_scaling
            gradient_max_norm = np.abs(x - plus(x, -grad)).max()
            step_is_successful = True
            strategy.step_accepted(relative_decrease)
        else:
            step_is_successful = False
            strategy.step_rejected(relative_decrease)
    return x, iteration, (cost0, x_cost)
