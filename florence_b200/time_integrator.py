"""Device-resident explicit central-difference loop (mechanics, lumped mass).

Replaces the time loop of ExplicitStructuralDynamicIntegrator.Solver
(Florence/TimeIntegrators/ExplicitStructuralDynamicIntegrator.py:28-244; helpers StructuralDynamicIntegrator.py:40-138,
:210-234).  All nDOF vectors (U0, U00, Eulerx, T, M, F_ext, Dirichlet mask) stay on the GPU; one step is
[fused node-reduction + update kernel] -> [blow-up test] -> [element internal-force kernel]; the host only sees the snapshots it
asks for (`save_frequency`) and the status word.

With an InterfaceExchange (one process per GPU, element-partitioned mesh) a step is
    update -> forces(all elements) -> gather+pack interface partial sums -> NCCL send/recv -> wait
           -> rank-ordered interface sums -> next update reads them for interface nodes
or, with overlap=True, forces(interface elements) -> pack -> send/recv posted -> forces(interior elements) while the messages are in
flight -> wait (same bits; measured slower on 4 and 8 ranks, see __init__), and the maxima of the blow-up test and the status word are agreed between the ranks, so every rank stops at the same increment.

Rigid-plane penalty contact (ExplicitPenaltyContactFormulation.AssembleTractions, called at :190-197) runs inside the same
kernels: pass `contact=` (an object with plane_normal, distance, kappa, contact_gap_tolerance) and the surface node list.

Out of scope, as in SURVEY.md H6: electro-mechanics (needs an implicit Poisson solve every step), consistent mass.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import backend

NAN_BIT, GROWTH_BIT = 1, 2
_I64_MIN = -(1 << 63)


class ExplicitStructuralDynamicIntegrator(object):

    def __init__(self, handle, material, M=None, rho=None, exchange=None, contact=None, surface_nodes=None, overlap=False,
                 check_growth=True):
        # overlap: evaluate the interface elements first and the interior ones while the messages travel (same bits either way,
        # tests/multigpu_check.py).  Off by default: on 4 and 8 ranks the split element kernel running beside the NCCL kernels of two
        # neighbours measured 0.1-3 ms per step SLOWER than evaluating everything and then exchanging (the point-to-point kernels
        # hold SM slots while they wait for a late neighbour), and on 2 ranks the two orders are within 0.5 % of each other
        # (profiles/r2_scaling_observed.json); two short tuning schemes picked the wrong order because the slowdown only develops
        # as the ranks drift apart.
        self.h = handle
        self.mat = material
        self.exchange = exchange
        self.overlap = overlap
        self.check_growth = check_growth
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
        self.status = torch.zeros(2, dtype=torch.int32, device=dev)
        self.keys = torch.full((2,), _I64_MIN, dtype=torch.int64, device=dev)
        self.last_status = (0, 0)      # (bits, increment of first detection) of the most recent step() call

    # kept under the round-1 name: tests and callers read it after a run
    @property
    def nan_flag(self):
        return self.status[:1]

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

    # -------------------------------------------------------------------------------------------- stepping
    def step(self, nsteps, increment, fext=None, fext_scale0=1.0, fext_scale_step=0.0, inc_dirichlet=None, incd_scale0=1.0,
             incd_scale_step=0.0):
        """Advance `nsteps` increments starting at `increment`.  Returns the status bits (0 = fine; NAN_BIT: a NaN appeared;
        GROWTH_BIT: abs(U.max()/(U0.max()+1e-14)) exceeded the reference's tolerance, :175-180); the same value on every rank.
        self.last_status also holds the increment of the first detection."""
        if self.exchange is None:
            st = self.h.explicit_steps(self.mat, self.dt, nsteps, increment, self.M, fext, self.fixed, inc_dirichlet, self.U0, self.U00,
                                       self.Eulerx, self.T, fext_scale0=fext_scale0, fext_scale_step=fext_scale_step,
                                       incd_scale0=incd_scale0, incd_scale_step=incd_scale_step, full_status=True)
            self.last_status = st
            self.status[0] = st[0]
            return st[0]
        return self._step_partitioned(nsteps, increment, fext, fext_scale0, fext_scale_step, inc_dirichlet, incd_scale0, incd_scale_step)

    def _force_and_exchange(self, x, nb):
        """Element forces of the whole local mesh + interface exchange; nb = number of leading interface elements whose forces are
        sent while the interior ones are evaluated (None / 0: evaluate everything, then exchange)."""
        h, ex, nv = self.h, self.exchange, self.ndim
        if nb is not None and 0 < nb < h.nelem:
            h.explicit_forces(x, self.mat, 0, nb)
            h.gather_pack_nodes(nv, ex.U, ex.own)
            ex.start(own_filled=True)
            h.explicit_forces(x, self.mat, nb, h.nelem)               # overlaps the messages
            ex.finish()
        else:
            h.explicit_forces(x, self.mat, 0, h.nelem)
            h.gather_pack_nodes(nv, ex.U, ex.own)
            ex.start(own_filled=True)
            ex.finish()

    def _step_partitioned(self, nsteps, increment, fext, fs0, fs1, inc_dirichlet, ds0, ds1):
        h, ex, nv = self.h, self.exchange, self.ndim
        x = self.Eulerx.view(self.nnode, nv)
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.status.zero_()
        for s in range(nsteps):
            inc = increment + s
            # the first update of a call consumes the caller-visible T (reduced, exchanged, with contact); later ones reduce the
            # per-element forces themselves and take interface nodes from the rank-ordered sums
            h.explicit_update(self.dt, fs0 + inc * fs1, self.M, fext, self.fixed, inc_dirichlet, self.T, self.U0, self.U00, self.Eulerx,
                              status=self.status, growth_keys=self.keys if self.check_growth else None, incd_scale=ds0 + inc * ds1,
                              use_element_forces=(s > 0), iface_slot=ex.iface_slot if s > 0 else None,
                              T_iface=ex.T_iface if s > 0 else None)
            if self.check_growth:
                if multi:
                    dist.all_reduce(self.keys, op=dist.ReduceOp.MAX)      # U.max(), U0.max() over the whole mesh
                h.explicit_check(self.keys, inc, self.status)
            self._force_and_exchange(x, self.exchange.part.n_interface_elements if self.overlap else None)
        if nsteps > 0:
            # caller-visible T: nodal reduction, interface nodes from the ordered sums, contact at the final geometry
            h.gather_nodes(nv, self.T)
            ex._scatter(self.T)
            if self.has_contact:
                h.assemble_contact(x, out=self.T, accumulate=True)
        st = self.status.clone()
        if multi:
            bits = st[:1].clone()
            first = torch.where(st[1:] > 0, st[1:], torch.full_like(st[1:], 1 << 30))
            dist.all_reduce(bits, op=dist.ReduceOp.MAX)
            dist.all_reduce(first, op=dist.ReduceOp.MIN)
            st = torch.cat([bits, first])
        b, f = int(st[0].item()), int(st[1].item())
        self.last_status = (b, f if b else 0)
        return b

    def blew_up(self):
        return bool(self.last_status[0]) or bool(self.status[0].item())

    def displacement(self):
        return (self.Eulerx - self.X).view(self.nnode, self.ndim)

    def run(self, nincrements, fext=None, ramp=True, inc_dirichlet=None, save_frequency=0, break_at_increment=None):
        """Increments 2 .. nincrements-1 as the reference's loop (:95); ramp loading F(inc) = F * inc / nincrements (:121-124).
        Snapshots (host numpy) are taken after every increment divisible by save_frequency (:165-166).  Stops, on every rank
        together, at a blow-up (:175-180) or after `break_at_increment` (:236-242)."""
        snaps = []
        inc = 2
        scale_step = 1.0 / nincrements if ramp else 0.0
        scale0 = 0.0 if ramp else 1.0 / max(nincrements - 1, 1)
        last = nincrements if break_at_increment in (None, -1) else min(nincrements, break_at_increment + 1)
        while inc < last:
            if save_frequency <= 0:
                n = last - inc
            else:
                n = min((save_frequency - inc % save_frequency) % save_frequency + 1, last - inc)   # ends on a multiple of save_frequency
            status = self.step(n, inc, fext, scale0, scale_step, inc_dirichlet)
            inc += n
            if status:
                print("Explicit solver blew up! Norm of incremental solution is too large")
                break
            if save_frequency > 0 and (inc - 1) % save_frequency == 0:
                snaps.append(self.displacement().cpu().numpy().copy())
        return snaps

    # -------------------------------------------------------------------------------------------- reference signature
    @classmethod
    def Solver(cls, function_spaces, formulation, solver, TractionForces, M, NeumannForces, NodalForces, Residual, mesh, TotalDisp,
               Eulerx, Eulerp, material, boundary_condition, fem_solver):
        """Drop-in for ExplicitStructuralDynamicIntegrator.Solver (ExplicitStructuralDynamicIntegrator.py:28-244), mechanics with
        lumped mass.  Same arguments; returns TotalDisp (nnode x nvar x nincrements) with the reference's save, blow-up
        (:175-180), break_at_increment (:236-242) and trimming (:244-251) rules.
        Ramp / constant loading given as single vectors (the common case, :103-108, :121-125) runs in chunks of
        `save_frequency` increments entirely on the device (the load and Dirichlet factors are formed in the kernel);
        step-wise tables (2-D applied_dirichlet / NeumannForces) upload one column per increment."""
        from . import assembly
        if formulation.fields != "mechanics":
            raise NotImplementedError("Explicit solver for {} is not available on this back end".format(formulation.fields))
        if getattr(fem_solver, "mass_type", "lumped") != "lumped":
            raise NotImplementedError("Only lumped mass is supported by the device-resident explicit loop")
        if getattr(fem_solver, "include_physical_damping", False):
            raise NotImplementedError("Damping is not included in the explicit solver")
        if getattr(boundary_condition, "has_step_wise_dirichlet_loading", False) or \
                getattr(boundary_condition, "has_step_wise_neumann_loading", False):
            raise NotImplementedError("Step-wise loading callbacks are not supported by the device-resident explicit loop")
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
        save_frequency = int(getattr(fem_solver, "save_frequency", 1))
        save_counter = 2 if save_frequency == 1 else 1
        nincr_last = float(LoadIncrement - 1) if LoadIncrement != 1 else 1
        applied = np.asarray(boundary_condition.applied_dirichlet, dtype=np.float64)
        cols_out = torch.as_tensor(np.asarray(boundary_condition.columns_out).astype(np.int64), device=dev)
        ramp = getattr(boundary_condition, "make_loading", "ramp") == "ramp"
        brk = getattr(fem_solver, "break_at_increment", -1)
        brk = None if brk in (None, -1) else int(brk)
        on_device = applied.ndim == 1 and NeumannForces.shape[1] == 1
        incd = torch.zeros_like(self.T)
        fext = None
        if on_device:
            incd[cols_out] = backend.to_device(applied, torch.float64, dev)
            fext = backend.to_device(np.ascontiguousarray(NeumannForces.ravel()), torch.float64, dev)
            s0, s1 = (0.0, 1.0 / LoadIncrement) if ramp else (1.0 / nincr_last, 0.0)       # factor(inc) = s0 + inc*s1

        def save_after(inc):
            return inc % save_frequency == 0 or (inc == LoadIncrement - 1 and save_counter < TotalDisp.shape[2])

        Increment = 2
        stopped = False
        while Increment < LoadIncrement and not stopped:
            if on_device:
                # run up to (and including) the next increment whose result has to be looked at on the host
                n = 1
                while Increment + n - 1 < LoadIncrement - 1 and not save_after(Increment + n - 1) and (brk is None or Increment + n - 1 != brk):
                    n += 1
                status = self.step(n, Increment, fext, s0, s1, incd, s0, s1)
            else:
                n = 1
                inc_col = applied[:, Increment - 1] if applied.ndim == 2 else \
                    (applied * (1. * Increment / LoadIncrement) if ramp else applied / nincr_last)          # :101-108
                incd.zero_()
                incd[cols_out] = backend.to_device(np.ascontiguousarray(inc_col), torch.float64, dev)
                if NeumannForces.shape[1] > 1:
                    fext = backend.to_device(np.ascontiguousarray(NeumannForces[:, Increment - 1]), torch.float64, dev)   # :119-120
                else:
                    fext = backend.to_device(NeumannForces.ravel() * ((1. * Increment / LoadIncrement) if ramp else 1.0 / nincr_last),
                                             torch.float64, dev)
                status = self.step(1, Increment, fext, 1.0, 0.0, incd, 1.0, 0.0)
            done = Increment + n - 1                  # last increment executed
            blown_at = self.last_status[1] if status else None
            # SAVE RESULTS (:165-172): the reference saves before it tests for blow-up, so the frame of the failing increment is
            # written when that increment is a saving one -- it is cut off again just below
            if (blown_at is None or blown_at == done) and save_after(done) and save_counter < TotalDisp.shape[2]:
                TotalDisp[:, :ndim, save_counter] = (self.Eulerx - self.X).view(nnode, ndim).cpu().numpy()
                save_counter += 1
            if status:
                print("Explicit solver blew up! Norm of incremental solution is too large")          # :177-180
                TotalDisp = TotalDisp[:, :, :blown_at]
                fem_solver.number_of_load_increments = blown_at
                stopped = True
            elif brk is not None and done == brk:                                                      # :236-242
                if brk < LoadIncrement - 1:
                    print("\nStopping at increment {} as specified\n\n".format(done))
                    TotalDisp = TotalDisp[:, :, :done]
                    fem_solver.number_of_load_increments = done
                stopped = True
            Increment = done + 1
        if save_frequency != 1:                                                                        # :244-251
            if TotalDisp.shape[2] > save_counter:
                TotalDisp = TotalDisp[:, :, :save_counter]
                fem_solver.number_of_load_increments = TotalDisp.shape[2]
            else:
                fem_solver.number_of_load_increments = save_counter
        if isinstance(Eulerx, np.ndarray):
            Eulerx[:, :] = self.Eulerx.view(nnode, ndim).cpu().numpy()
        return TotalDisp
