"""Device-side adaptive Dormand-Prince RK45 for the probability-flow ODE sampler.

The reference (sampling.py:342-392) calls scipy.integrate.solve_ivp: the state lives on the HOST as a float64 numpy
vector and every right-hand side moves it host -> GPU -> host (6 times per step).  This solver applies the same method
with the same control law -- scipy/integrate/_ivp/rk.py `RK45` + `RungeKutta._step_impl` + `select_initial_step`: one
global RMS error norm over the whole batch, safety 0.9, factors in [0.2, 10], first-same-as-last -- but keeps the float64
state y, the seven fp32 stage derivatives K and the stage arithmetic on the GPU (csrc/next_rows.cu rd_rk45_*).  Per
attempted step the host reads ONE double (the squared error norm) to take scipy's accept / reject decision.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib, stream_ptr

# Dormand-Prince coefficients (scipy.integrate._ivp.rk.RK45)
C_NODES = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
A_ROWS = [np.array(r, dtype=np.float64) for r in (
    [], [1 / 5], [3 / 40, 9 / 40], [44 / 45, -56 / 15, 32 / 9], [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656])]
B_ROW = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84], dtype=np.float64)
E_ROW = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40], dtype=np.float64)
SAFETY, MIN_FACTOR, MAX_FACTOR = 0.9, 0.2, 10.0
ERROR_ESTIMATOR_ORDER = 4
ERROR_EXPONENT = -1.0 / (ERROR_ESTIMATOR_ORDER + 1)


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class DeviceRK45:
    """solve_ivp(fun, (t0, t_bound), y0, method='RK45', rtol, atol) with `fun(t, x_fp32_device) -> fp32 device tensor`."""

    def __init__(self, fun, t0: float, y0: torch.Tensor, t_bound: float, rtol: float = 1e-5, atol: float = 1e-5):
        if not y0.is_cuda:
            raise RuntimeError("DeviceRK45 runs on CUDA tensors only (no CPU path)")
        self.fun, self.shape, self.dev = fun, tuple(y0.shape), y0.device
        self.n = y0.numel()
        self.rtol, self.atol = float(rtol), float(atol)
        self.t, self.t_bound = float(t0), float(t_bound)
        self.direction = float(np.sign(t_bound - t0)) if t_bound != t0 else 1.0
        self.y = y0.detach().to(torch.float64).reshape(-1).contiguous()     # float64 like scipy's state
        self.y_new = torch.empty_like(self.y)
        self.x32 = torch.empty(self.shape, dtype=torch.float32, device=self.dev)
        self.K = torch.zeros((7, self.n), dtype=torch.float32, device=self.dev)
        self.partial = torch.empty((148 * 8,), dtype=torch.float64, device=self.dev)
        self.sumsq = torch.zeros((1,), dtype=torch.float64, device=self.dev)
        self.nfev = 0
        self.status = "running"
        self._eval_into(0, self.t, None, 0, 0.0)                              # f0 = fun(t0, y0)
        self.h_abs = self._select_initial_step()

    # -- one right-hand side: stage vector on the device -> network -> K[slot]
    def _eval_into(self, slot: int, t: float, a, s: int, h: float, y_out=None):
        st = stream_ptr(self.dev)
        arr = np.zeros(6, dtype=np.float64) if a is None else np.ascontiguousarray(a, dtype=np.float64)
        check(lib().rd_rk45_stage_f64(self.y.data_ptr(), self.K.data_ptr(), self.n, s, _dptr(arr), h,
                                      None if y_out is None else y_out.data_ptr(), self.x32.data_ptr(), st), "rd_rk45_stage_f64")
        f = self.fun(t, self.x32)
        self.K[slot].copy_(f.reshape(-1))
        self.nfev += 1

    def _rms(self, v: torch.Tensor) -> float:
        return float(torch.linalg.vector_norm(v) / np.sqrt(v.numel()))

    def _select_initial_step(self) -> float:
        """scipy.integrate._ivp.common.select_initial_step (once per solve: a few fp64 reductions on the device)."""
        if self.n == 0:
            return float("inf")
        interval = abs(self.t_bound - self.t)
        if interval == 0.0:
            return 0.0
        y0, f0 = self.y, self.K[0].to(torch.float64)
        scale = self.atol + y0.abs() * self.rtol
        d0, d1 = self._rms(y0 / scale), self._rms(f0 / scale)
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        h0 = min(h0, interval)
        # f1 = fun(t0 + h0 * direction, y0 + h0 * direction * f0), evaluated into the scratch slot 1
        self._eval_into(1, self.t + h0 * self.direction, np.array([1.0]), 1, h0 * self.direction)
        f1 = self.K[1].to(torch.float64)
        d2 = self._rms((f1 - f0) / scale) / h0
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = max(1e-6, h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** (1.0 / (ERROR_ESTIMATOR_ORDER + 1))
        return min(100 * h0, h1, interval)

    def step(self) -> bool:
        """RungeKutta._step_impl: attempt steps until one is accepted (True) or the step size underflows (False)."""
        t, st = self.t, stream_ptr(self.dev)
        min_step = 10 * abs(np.nextafter(t, self.direction * np.inf) - t)
        h_abs = max(self.h_abs, min_step)   # (max_step is infinite: solve_ivp's default)
        rejected = False
        while True:
            if h_abs < min_step:
                self.status = "failed"
                return False
            h = h_abs * self.direction
            t_new = t + h
            if self.direction * (t_new - self.t_bound) > 0:
                t_new = self.t_bound
            h = t_new - t
            h_abs = abs(h)
            for s in range(1, 6):                                   # K_1 .. K_5 (K_0 = f(t, y) is carried over)
                self._eval_into(s, t + C_NODES[s] * h, A_ROWS[s], s, h)
            self._eval_into(6, t + h, B_ROW, 6, h, y_out=self.y_new)  # y_new and f_new = K_6
            check(lib().rd_rk45_error_f64(self.y.data_ptr(), self.y_new.data_ptr(), self.K.data_ptr(), self.n, _dptr(E_ROW), h, self.atol,
                                          self.rtol, self.partial.data_ptr(), self.partial.numel(), self.sumsq.data_ptr(), st),
                  "rd_rk45_error_f64")
            err = float(np.sqrt(float(self.sumsq.item()) / self.n))   # the one host read of the step
            if err < 1:
                factor = MAX_FACTOR if err == 0 else min(MAX_FACTOR, SAFETY * err ** ERROR_EXPONENT)
                if rejected:
                    factor = min(1.0, factor)
                h_abs *= factor
                break
            h_abs *= max(MIN_FACTOR, SAFETY * err ** ERROR_EXPONENT)
            rejected = True
        self.t = t_new
        self.y, self.y_new = self.y_new, self.y
        self.K[0].copy_(self.K[6])                                    # first same as last
        self.h_abs = h_abs
        if self.direction * (self.t - self.t_bound) >= 0:
            self.status = "finished"
        return True

    def solve(self) -> torch.Tensor:
        while self.status == "running":
            if self.n == 0 or self.t == self.t_bound:
                self.status = "finished"
                break
            if not self.step():
                raise RuntimeError("DeviceRK45: required step size is less than spacing between numbers")
        return self.y.reshape(self.shape)
