"""Device-resident explicit central-difference loop (mechanics, lumped mass).

Replaces the time loop of ExplicitStructuralDynamicIntegrator.Solver
(Florence/TimeIntegrators/ExplicitStructuralDynamicIntegrator.py:28-244; helpers StructuralDynamicIntegrator.py:40-138,
:210-234).  All nDOF vectors (U0, U00, Eulerx, T, M, F_ext, Dirichlet mask) stay on the GPU; one step is
[fused node-reduction + update kernel] -> [element internal-force kernel]; the host only sees the snapshots it asks for
(`save_frequency`) and the blow-up flag.  With an InterfaceExchange the loop becomes force -> exchange -> update per step.

Rigid-plane penalty contact (ExplicitPenaltyContactFormulation.AssembleTractions, called at :190-197) runs inside the same
kernels: pass `contact=` (an object with plane_normal, distance, kappa, contact_gap_tolerance) and the surface node list.

Out of scope, as in SURVEY.md H6: electro-mechanics (needs an implicit Poisson solve every step), consistent mass.
"""
import numpy as np
import torch

from . import backend


class ExplicitStructuralDynamicIntegrator(object):

    def __init__(self, handle, material, M=None, rho=None, exchange=None, contact=None, surface_nodes=None):
        self.h = handle
        self.mat = material
        self.exchange = exchange
        self.has_contact = contact is not None
        if self.has_contact:
            # fem_solver.contact_formulation of the reference (FEMSolver.py:215-218); surface_nodes = np.unique(mesh.faces)
            handle.set_contact(surface_nodes, contact.plane_normal, contact.distance, contact.kappa,
                               getattr(contact, "contact_gap_tolerance", 1e-6))
        else:
            handle.set_contact(None, None, 0.0, 0.0)
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
        """TractionForces of the loop: AssembleExplicit (+ interface sum) (+ contact tractions, :190-197)."""
        T = self.h.assemble_explicit(Eulerx, None, self.mat, 0, out=out)
        if self.exchange is not None:
            self.exchange(T)
        if self.has_contact:
            self.h.assemble_contact(Eulerx, out=T, accumulate=True)
        return T

    def initialise(self, X, fext0, fixed_mask, dt):
        """Start-up of ExplicitStructuralDynamicIntegrator.py:57-81 with U0 = V0 = 0:
        A0 = invM (F_ext(0) - T), U00 = dt^2/2 A0, fixed dofs zeroed."""
        dev = self.h.device
        self.X = backend.to_device(X, torch.float64, dev).reshape(-1)
        self.fixed = backend.to_device(fixed_mask, torch.uint8, dev).reshape(-1)
        self.Eulerx = self.X.clone()
        # the reference's first TractionForces comes from FEMSolver's initial Assemble, before any contact check
        self.T = self.h.assemble_explicit(self.Eulerx.view(self.nnode, self.ndim), None, self.mat, 0)
        if self.exchange is not None:
            self.exchange(self.T)
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


    # -------------------------------------------------------------------------------------------- reference signature
    @classmethod
    def Solver(cls, function_spaces, formulation, solver, TractionForces, M, NeumannForces, NodalForces, Residual, mesh, TotalDisp,
               Eulerx, Eulerp, material, boundary_condition, fem_solver):
        """Drop-in for ExplicitStructuralDynamicIntegrator.Solver (ExplicitStructuralDynamicIntegrator.py:28-244), mechanics with
        lumped mass.  Same arguments; returns TotalDisp (nnode x nvar x nincrements) with the reference's save rules.
        The loop runs on the device; per increment the host only selects the load / Dirichlet columns."""
        from . import assembly
        if formulation.fields != "mechanics":
            raise NotImplementedError("Explicit solver for {} is not available on this back end".format(formulation.fields))
        if getattr(fem_solver, "mass_type", "lumped") != "lumped":
            raise NotImplementedError("Only lumped mass is supported by the device-resident explicit loop")
        if getattr(fem_solver, "include_physical_damping", False):
            raise NotImplementedError("Damping is not included in the explicit solver")
        fspace = function_spaces[1] if len(function_spaces) > 1 else function_spaces[0]
        h = assembly.get_handle(mesh, fspace)
        mat = assembly._material_struct(material)
        dev = h.device
        ndim, nnode = formulation.ndim, mesh.points.shape[0]
        LoadIncrement = fem_solver.number_of_load_increments
        dt = fem_solver.total_time / LoadIncrement
        contact = getattr(fem_solver, "contact_formulation", None) if getattr(fem_solver, "has_contact", False) else None
        surface = None
        if contact is not None:
            surface = np.unique(mesh.edges if ndim == 2 else mesh.faces).astype(np.int64)      # ExplicitPenaltyContactFormulation.py:157-161
        self = cls(h, mat, M=np.asarray(M, dtype=np.float64).ravel(), contact=contact, surface_nodes=surface)
        NeumannForces = np.asarray(NeumannForces, dtype=np.float64)
        if NeumannForces.ndim == 1:
            NeumannForces = NeumannForces[:, None]
        fixed = np.zeros(nnode * ndim, dtype=np.uint8)
        fixed[np.asarray(boundary_condition.columns_out)] = 1
        self.X = backend.to_device(mesh.points, torch.float64, dev).reshape(-1)
        self.fixed = backend.to_device(fixed, torch.uint8, dev)
        self.Eulerx = backend.to_device(Eulerx, torch.float64, dev).reshape(-1).clone()
        self.T = backend.to_device(np.asarray(TractionForces, dtype=np.float64).ravel(), torch.float64, dev).clone()
        # start-up, :57-81 (zero initial displacement and velocity)
        A0 = (backend.to_device(np.ascontiguousarray(NeumannForces[:, 0]), torch.float64, dev) - self.T) / self.M
        self.U0 = torch.zeros_like(self.T)
        self.U00 = (dt ** 2 / 2.) * A0
        self.U00[self.fixed.bool()] = 0.0
        self.dt = dt
        TotalDisp[:, :ndim, 0] = self.U00.view(nnode, ndim).cpu().numpy()
        save_frequency = getattr(fem_solver, "save_frequency", 1)
        save_counter = 2 if save_frequency == 1 else 1
        nincr_last = float(LoadIncrement - 1) if LoadIncrement != 1 else 1
        applied = np.asarray(boundary_condition.applied_dirichlet, dtype=np.float64)
        cols_out = torch.as_tensor(np.asarray(boundary_condition.columns_out).astype(np.int64), device=dev)
        incd = torch.zeros_like(self.T)
        ramp = getattr(boundary_condition, "make_loading", "ramp") == "ramp"
        for Increment in range(2, LoadIncrement):
            # incremental Dirichlet / Neumann data, :101-128
            if applied.ndim == 2:
                inc_col = applied[:, Increment - 1]
            else:
                inc_col = applied * (1. * Increment / LoadIncrement) if ramp else applied / nincr_last
            incd.zero_()
            incd[cols_out] = backend.to_device(np.ascontiguousarray(inc_col), torch.float64, dev)
            if NeumannForces.shape[1] > 1:
                fext = backend.to_device(np.ascontiguousarray(NeumannForces[:, Increment - 1]), torch.float64, dev)
            else:
                fext = backend.to_device(NeumannForces.ravel() * ((1. * Increment / LoadIncrement) if ramp else 1.0 / nincr_last), torch.float64, dev)
            status = h.explicit_steps(mat, dt, 1, Increment, self.M, fext, self.fixed, incd, self.U0, self.U00, self.Eulerx, self.T,
                                      fext_scale0=1.0, fext_scale_step=0.0)
            if Increment % save_frequency == 0 or (Increment == LoadIncrement - 1 and save_counter < TotalDisp.shape[2]):
                TotalDisp[:, :ndim, save_counter] = (self.Eulerx - self.X).view(nnode, ndim).cpu().numpy()
                save_counter += 1
            if status:
                print("Explicit solver blew up! Norm of incremental solution is too large")
                TotalDisp = TotalDisp[:, :, :Increment]
                fem_solver.number_of_load_increments = Increment
                break
        if isinstance(Eulerx, np.ndarray):
            Eulerx[:, :] = self.Eulerx.view(nnode, ndim).cpu().numpy()
        return TotalDisp
