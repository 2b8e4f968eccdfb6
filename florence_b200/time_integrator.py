"""Device-resident explicit central-difference loop (mechanics, lumped mass).

Replaces the time loop of ExplicitStructuralDynamicIntegrator.Solver
(Florence/TimeIntegrators/ExplicitStructuralDynamicIntegrator.py:28-244; helpers StructuralDynamicIntegrator.py:40-138,
:210-234).  All nDOF vectors (U0, U00, Eulerx, T, M, F_ext, Dirichlet mask) stay on the GPU; one step is
[fused node-reduction + update kernel] -> [element internal-force kernel]; the host only sees the snapshots it asks for
(`save_frequency`) and the blow-up flag.  With an InterfaceExchange the loop becomes force -> exchange -> update per step.

Out of scope, as in SURVEY.md H6: electro-mechanics (needs an implicit Poisson solve every step), consistent mass, contact.
"""
import numpy as np
import torch

from . import backend


class ExplicitStructuralDynamicIntegrator(object):

    def __init__(self, handle, material, M=None, rho=None, exchange=None):
        self.h = handle
        self.mat = material
        self.exchange = exchange
        self.nnode, self.ndim = handle.nnode, handle.ndim
        dev = handle.device
        if M is None:
            # __TotalConstantMassIntegrand__ (lumped), summed over ranks on the interface
            M = handle.assemble_mass(material.rho if rho is None else rho, self.ndim, "lumped")
            if exchange is not None:
                exchange(M)
        self.M = backend.to_device(M, torch.float64, dev).reshape(-1)
        self.X = None
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)

    def internal_force(self, Eulerx, out=None):
        T = self.h.assemble_explicit(Eulerx, None, self.mat, 0, out=out)
        if self.exchange is not None:
            self.exchange(T)
        return T

    def initialise(self, X, fext0, fixed_mask, dt):
        """Start-up of ExplicitStructuralDynamicIntegrator.py:57-81 with U0 = V0 = 0:
        A0 = invM (F_ext(0) - T), U00 = dt^2/2 A0, fixed dofs zeroed."""
        dev = self.h.device
        self.X = backend.to_device(X, torch.float64, dev).reshape(-1)
        self.fixed = backend.to_device(fixed_mask, torch.uint8, dev).reshape(-1)
        self.Eulerx = self.X.clone()
        self.T = self.internal_force(self.Eulerx.view(self.nnode, self.ndim))
        f0 = torch.zeros_like(self.T) if fext0 is None else backend.to_device(fext0, torch.float64, dev).reshape(-1)
        A0 = (f0 - self.T) / self.M
        self.U0 = torch.zeros_like(self.T)
        self.U00 = (dt ** 2 / 2.) * A0
        self.U00[self.fixed.bool()] = 0.0
        self.dt = dt
        return self

    def step(self, nsteps, increment, fext=None, fext_scale0=1.0, fext_scale_step=0.0, inc_dirichlet=None):
        """Advance `nsteps` increments.  Returns 1 if the solution blew up (NaN), else 0 (:175-180)."""
        if self.exchange is None:
            return self.h.explicit_steps(self.mat, self.dt, nsteps, increment, self.M, fext, self.fixed, inc_dirichlet, self.U0, self.U00,
                                         self.Eulerx, self.T, fext_scale0=fext_scale0, fext_scale_step=fext_scale_step)
        # multi-GPU: update (needs the exchanged T) -> internal force -> interface exchange
        for s in range(nsteps):
            fs = fext_scale0 + (increment + s) * fext_scale_step
            self.h.explicit_update(self.dt, fs, self.M, fext, self.fixed, inc_dirichlet, self.T, self.U0, self.U00, self.Eulerx, self.nan_flag)
            self.internal_force(self.Eulerx.view(self.nnode, self.ndim), out=self.T)
        return 0

    def blew_up(self):
        return bool(self.nan_flag.item())

    def displacement(self):
        return (self.Eulerx - self.X).view(self.nnode, self.ndim)

    def run(self, nincrements, fext=None, ramp=True, inc_dirichlet=None, save_frequency=0):
        """Increments 2 .. nincrements-1 as the reference's loop (:95); ramp loading F(inc) = F * inc / nincrements (:121-124).
        Returns the list of saved displacement snapshots (host numpy) when save_frequency > 0."""
        snaps = []
        inc = 2
        scale_step = 1.0 / nincrements if ramp else 0.0
        scale0 = 0.0 if ramp else 1.0 / max(nincrements - 1, 1)
        while inc < nincrements:
            n = nincrements - inc if save_frequency <= 0 else min(save_frequency - inc % save_frequency, nincrements - inc)
            status = self.step(n, inc, fext, scale0, scale_step, inc_dirichlet)
            inc += n
            if save_frequency > 0:
                snaps.append(self.displacement().cpu().numpy().copy())
            if status:
                print("Explicit solver blew up! Norm of incremental solution is too large")
                break
        return snaps
