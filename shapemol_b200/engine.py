"""Host-side driver of the CUDA library: packs weights, owns workspaces, runs the step loop.

PyTorch is plumbing here (device memory, streams, the RNG stream the reference uses); all arithmetic
of the hot path is enqueued through the C ABI (shapemol_b200/_lib.py).  There is no eager fallback.
"""
import ctypes as C

import torch

from . import _lib

SCHEDULE_TABLES = ('posterior_mean_c0_coef', 'posterior_mean_ct_coef', 'posterior_logvar', 'log_alphas_v',
                   'log_one_minus_alphas_v', 'log_alphas_cumprod_v', 'log_one_minus_alphas_cumprod_v')


class BatchDesc:
    """Ragged batch descriptor (mol_ptr / atom_mol on the device) built once per batch."""

    def __init__(self, batch_ligand, n_mols=None):
        dev = batch_ligand.device
        if dev.type != 'cuda':
            raise _lib.SmbError('shapemol_b200 runs on CUDA devices only (got %s); there is no CPU path' % dev)
        self.device = dev
        self.n_atoms = int(batch_ligand.numel())
        if n_mols is None:
            n_mols = int(batch_ligand.max().item()) + 1 if self.n_atoms else 0   # same sync as the reference (:290)
        self.n_mols = n_mols
        counts = torch.bincount(batch_ligand, minlength=max(n_mols, 1))[:max(n_mols, 1)]
        self.max_atoms = int(counts.max().item()) if self.n_atoms else 0
        if self.max_atoms > _lib.SMB_MAX_ATOMS_PER_MOL:
            raise _lib.SmbError('molecule with %d atoms exceeds SMB_MAX_ATOMS_PER_MOL=%d' % (self.max_atoms, _lib.SMB_MAX_ATOMS_PER_MOL))
        ptr = torch.zeros(n_mols + 1, dtype=torch.int32, device=dev)
        if n_mols:
            ptr[1:] = torch.cumsum(counts[:n_mols], 0).to(torch.int32)
        self.mol_ptr = ptr
        self.atom_mol = batch_ligand.to(torch.int32).contiguous()
        if self.n_atoms > 1 and bool((self.atom_mol[1:] < self.atom_mol[:-1]).any().item()):
            raise _lib.SmbError('batch_ligand must be sorted (atoms of a molecule contiguous)')
        self.c = _lib.Batch(self.n_atoms, self.n_mols, self.max_atoms, self.mol_ptr.data_ptr(), self.atom_mol.data_ptr())


class DenoiseEngine:
    """One engine per ScorePosNet3D module instance."""

    def __init__(self, module, precision='bf16x3'):
        self.module = module
        self.lib = _lib.load()
        self.set_precision(precision)
        self._packed = None
        self._packed_key = None
        self._ws = None
        self._static_key = None  # (batch, shape, workspace, weights) whose step-independent quantities the workspace holds
        self._names = None

    # ---- configuration ---------------------------------------------------------------------
    def set_precision(self, precision):
        if precision not in _lib.PRECISIONS:
            raise ValueError('precision must be one of %s' % sorted(_lib.PRECISIONS))
        self.precision = precision
        self._packed_key = None

    def dims(self):
        m = self.module
        rn = m.refine_net
        return _lib.ModelDims(hidden=m.hidden_dim, heads=rn.n_heads, layers=rn.num_layers, k=rn.k, classes=m.num_classes,
                              time_dim=m.time_emb_dim, timesteps=m.num_timesteps, precision=_lib.PRECISIONS[self.precision])

    # ---- weights ---------------------------------------------------------------------------
    def _param_names(self, dims):
        if self._names is None:
            n = self.lib.smb_param_count(C.byref(dims))
            if n <= 0:
                _lib.check(n if n < 0 else -1, 'smb_param_count')
            self._names = [self.lib.smb_param_name(C.byref(dims), i).decode() for i in range(n)]
        return self._names

    def packed_weights(self, device):
        dims = self.dims()
        names = self._param_names(dims)
        sd = self.module.state_dict(keep_vars=True)
        key = (self.precision, str(device)) + tuple((sd[n].data_ptr(), sd[n]._version) for n in names)
        if key != self._packed_key:
            host = [sd[n].detach().to('cpu', torch.float32).contiguous() for n in names]
            arr = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
            nbytes = self.lib.smb_packed_weights_bytes(C.byref(dims))
            blob = torch.empty(nbytes, dtype=torch.uint8)
            _lib.check(self.lib.smb_pack_weights(C.byref(dims), arr, len(host), blob.data_ptr(), nbytes), 'smb_pack_weights')
            self._packed = blob.to(device)
            self._packed_key = key
        return self._packed

    def workspace(self, dims, n_atoms, n_mols, device, owner=None):
        """Per-step scratch.  `owner` (a Sampler / HostStepper) gets a workspace of its own: a captured CUDA graph bakes the
        workspace pointer in, so a later, larger batch evaluated through the same engine must not reallocate it."""
        need = self.lib.smb_workspace_bytes(C.byref(dims), n_atoms, n_mols)
        if owner is not None:
            ws = getattr(owner, '_ws', None)
            if ws is None or ws.numel() < need or ws.device != device:
                if getattr(owner, 'graph', None) is not None or getattr(owner, '_graph', None) is not None:
                    raise _lib.SmbError('workspace of a captured step graph cannot grow; build a new Sampler / HostStepper')
                ws = torch.empty(need, dtype=torch.uint8, device=device)
                owner._ws = ws
            return ws
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def _bn_layers(self):
        return [blk.h2x_layers[0].shape_linear.batchnorm.bn for blk in self.module.refine_net.base_block]

    # ---- one network evaluation --------------------------------------------------------------
    def forward(self, pos, v_i32, bd, shape, t_i32, pred_pos, pred_h, pred_v, h0=None, nbr=None, training=None, prof=None,
                reuse_static=False, owner=None):
        if bd.n_atoms == 0:
            return                      # empty batch (e.g. a rank whose shard holds no molecule): nothing to enqueue
        dims = self.dims()
        dev = pos.device
        blob = self.packed_weights(dev)
        ws = self.workspace(dims, bd.n_atoms, bd.n_mols, dev, owner)
        io = _lib.ForwardIO()
        io.pos, io.v, io.shape, io.t = pos.data_ptr(), v_i32.data_ptr(), shape.data_ptr(), t_i32.data_ptr()
        io.pred_pos, io.pred_h, io.pred_v = pred_pos.data_ptr(), pred_h.data_ptr(), pred_v.data_ptr()
        io.h0, io.nbr = _lib.ptr(h0), _lib.ptr(nbr)
        for l, bn in enumerate(self._bn_layers()):
            for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var):
                if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                    raise _lib.SmbError('BatchNorm tensors must be contiguous fp32 on %s' % dev)
            io.bn_weight[l], io.bn_bias[l] = bn.weight.data_ptr(), bn.bias.data_ptr()
            io.bn_running_mean[l], io.bn_running_var[l] = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
            io.bn_num_batches_tracked[l] = bn.num_batches_tracked.data_ptr()
        io.training = int(self.module.training if training is None else training)
        # step-independent workspace contents (invariant shape embedding, VN shape maps, tile list) are kept only if the
        # previous full evaluation used the same batch descriptor, shape tensor, workspace and packed weights
        # (the shape tensor's version counter catches in-place updates of the condition latents)
        key = (id(bd), shape.data_ptr(), shape._version, ws.data_ptr(), blob.data_ptr())
        reuse = bool(reuse_static) and key == self._static_key
        self._static_key = key
        io.reuse_static = int(reuse)
        if prof is not None:      # (kernel class name, [torch.cuda.Event pairs, already created])
            cls, events = prof
            handles = (C.c_void_p * len(events))(*[ev.cuda_event for ev in events])
            io.prof_kernel, io.prof_capacity, io.prof_events = _lib.PROF[cls], len(events) // 2, handles
        with torch.cuda.device(dev):     # the launchers configure / launch on the CURRENT device
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(self.lib.smb_forward(C.byref(dims), blob.data_ptr(), C.byref(bd.c), C.byref(io), ws.data_ptr(), ws.numel(), stream),
                       'smb_forward')

    def type_head(self, h, bd, logits):
        dims = self.dims()
        blob = self.packed_weights(h.device)
        stream = torch.cuda.current_stream(h.device).cuda_stream
        _lib.check(self.lib.smb_type_head(C.byref(dims), blob.data_ptr(), C.byref(bd.c), h.data_ptr(), logits.data_ptr(), stream),
                   'smb_type_head')

    # ---- one reverse-diffusion update ----------------------------------------------------------
    def posterior(self, bd, pred_pos, pred_v, t_i32, pos, v_i32, noise_pos=None, noise_u=None, log_v0=None, log_post=None,
                  seed=0, atom_offset=0):
        dims = self.dims()
        m = self.module
        io = _lib.PosteriorIO()
        io.pred_pos, io.pred_v, io.t, io.pos, io.v = pred_pos.data_ptr(), pred_v.data_ptr(), t_i32.data_ptr(), pos.data_ptr(), v_i32.data_ptr()
        io.noise_pos, io.noise_u, io.log_v0, io.log_post = _lib.ptr(noise_pos), _lib.ptr(noise_u), _lib.ptr(log_v0), _lib.ptr(log_post)
        io.seed, io.atom_offset = int(seed) & (2 ** 64 - 1), int(atom_offset)
        for name in SCHEDULE_TABLES:
            tab = getattr(m, name)
            if tab.device != pos.device:
                raise _lib.SmbError('schedule table %s is not on %s' % (name, pos.device))
            setattr(io, name, tab.data_ptr())
        stream = torch.cuda.current_stream(pos.device).cuda_stream
        _lib.check(self.lib.smb_posterior_step(C.byref(dims), C.byref(bd.c), C.byref(io), stream), 'smb_posterior_step')

    # ---- point-cloud shape guidance on the predicted x0 (reference :582-591, :699-740) ----------------
    def guidance(self, bd, pos, cloud_f64, radius, ratio=0.2, t_i32=None, grad_step=0, step=0, u=None, cloud_ptr=None, seed=0,
                 atom_offset=0):
        if cloud_f64.dtype != torch.float64 or cloud_f64.device != pos.device or not cloud_f64.is_contiguous():
            raise _lib.SmbError('the condition point cloud must be a contiguous float64 [M,3] tensor on %s' % pos.device)
        io = _lib.GuidanceIO()
        io.pos, io.cloud, io.cloud_ptr, io.n_cloud = pos.data_ptr(), cloud_f64.data_ptr(), _lib.ptr(cloud_ptr), int(cloud_f64.shape[0])
        io.t, io.grad_step, io.step = _lib.ptr(t_i32), int(grad_step), int(step)
        io.radius, io.ratio, io.u = float(radius), float(ratio), _lib.ptr(u)
        io.seed, io.atom_offset = int(seed) & (2 ** 64 - 1), int(atom_offset)
        stream = torch.cuda.current_stream(pos.device).cuda_stream
        _lib.check(self.lib.smb_pointcloud_guidance(C.byref(bd.c), C.byref(io), stream), 'smb_pointcloud_guidance')

    # ---- alignment-free shape Tanimoto of every molecule against its reference centres (get_ROCS) ------
    def shape_tanimoto(self, bd, pos, ref_f64, ref_ptr=None, prefactor=0.8, alpha=0.81):
        import math
        import numpy as np
        if ref_f64.dtype != torch.float64 or ref_f64.device != pos.device or not ref_f64.is_contiguous():
            raise _lib.SmbError('reference centres must be a contiguous float64 [R,3] tensor on %s' % pos.device)
        # the reference's per-atom constants are float32 (torch.ones(n) * x, utils/evaluation/shaep_utils.py:76-79)
        a, p = np.float32(alpha), np.float32(prefactor)
        k = float(np.float32(np.float32(a * a) / np.float32(a + a)))
        coef = float(np.float32(np.float32(math.pi ** 1.5) * np.float32(p * p)))
        den = float(np.float32(np.float32(a + a) ** np.float32(1.5)))
        out = torch.empty(bd.n_mols, dtype=torch.float64, device=pos.device)
        stream = torch.cuda.current_stream(pos.device).cuda_stream
        _lib.check(self.lib.smb_shape_tanimoto(C.byref(bd.c), pos.data_ptr(), ref_f64.data_ptr(), _lib.ptr(ref_ptr), int(ref_f64.shape[0]),
                                               k, coef, den, out.data_ptr(), stream), 'smb_shape_tanimoto')
        return out

    # ---- stability check of every molecule (check_stability, utils/evaluation/analyze.py:264-297) -----------
    def check_stability(self, bd, pos, atomic_numbers, hs=False):
        """-> (molecule_stable [B] bool, stable_atoms [B] int32, nr_bonds [N] int32), all on the device."""
        from . import chem_tables as ct
        dev = pos.device
        elem = ct.element_index(atomic_numbers).to(torch.int32).to(dev).contiguous()
        thr = ct.thresholds().to(dev).contiguous()
        allowed = torch.tensor(ct.ALLOWED_BONDS, dtype=torch.int32, device=dev)
        nr = torch.empty(bd.n_atoms, dtype=torch.int32, device=dev)
        st_atoms = torch.empty(bd.n_mols, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self.lib.smb_check_stability(C.byref(bd.c), pos.data_ptr(), elem.data_ptr(), thr.data_ptr(), allowed.data_ptr(),
                                                len(ct.ELEMENTS), int(bool(hs)), nr.data_ptr(), st_atoms.data_ptr(), stream), 'smb_check_stability')
        sizes = (bd.mol_ptr[1:] - bd.mol_ptr[:-1]).to(torch.int32)
        return st_atoms == sizes, st_atoms, nr

    def decrement_t(self, t_i32):
        stream = torch.cuda.current_stream(t_i32.device).cuda_stream
        _lib.check(self.lib.smb_decrement_t(t_i32.data_ptr(), t_i32.numel(), stream), 'smb_decrement_t')

    def knn_graph(self, x, bd, k):
        nbr = torch.empty(bd.n_atoms, k + 1, dtype=torch.int32, device=x.device)
        deg = torch.empty(bd.n_atoms, dtype=torch.int32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self.lib.smb_knn_graph(x.data_ptr(), C.byref(bd.c), k, nbr.data_ptr(), deg.data_ptr(), stream), 'smb_knn_graph')
        return nbr, deg


class Sampler:
    """Reverse-diffusion loop (ScorePosNet3D.sample_diffusion default branch) with persistent device
    state, no host synchronisation inside the loop and the step captured in a CUDA graph.

    A Sampler owns its workspace (the captured graph bakes its address in), so other batches may be evaluated through the
    same engine while it exists.

    noise = 'torch'  : per step torch.randn_like(pos) then torch.rand(N, C) from the device's default
                       generator -- the reference's draw order, so torch.manual_seed(s) reproduces it;
            'philox' : in-kernel Philox4x32 keyed by (seed, global atom index, t)   (throughput mode);
            callable : noise(step) -> (randn [N,3], rand [N,C])                    (parity tests).

    Trajectories (keep_traj=True; reference :671-681): the four lists the reference moves to the CPU every step (pos, v,
    v0, vt) are recorded into a double-buffered device ring of `traj_chunk` steps and drained to host tensors [S,N,.] on a
    side stream, one D2H per list per chunk, while the next chunk's steps run; the two lists the reference keeps on the
    device (pos_cond, v_cond) are device tensors [S,N,.].
    """

    HOST_KEYS = ('pos', 'v', 'v0', 'vt')

    def __init__(self, engine, init_pos, init_v, batch_ligand, shape, num_steps=None, noise='torch', seed=0, atom_offset=0,
                 keep_traj=True, use_graph=True, n_mols=None, guidance=None, traj_chunk=None):
        m = engine.module
        self.e = engine
        dev = init_pos.device
        self.bd = BatchDesc(batch_ligand, n_mols)
        N, B, Cn, H = self.bd.n_atoms, self.bd.n_mols, m.num_classes, m.hidden_dim
        T = m.num_timesteps
        self.num_steps = T if num_steps is None else int(num_steps)
        self.pos = init_pos.detach().to(torch.float32).clone().contiguous()
        self.v = init_v.detach().to(torch.int32).contiguous().clone()
        self.shape = shape.detach().to(torch.float32).reshape(B, -1, 3).contiguous() if B else torch.zeros(0, 32, 3, device=dev)
        self.t = torch.full((max(B, 1),), T - 1, dtype=torch.int32, device=dev)
        self.pred_pos = torch.empty(N, 3, device=dev)
        self.pred_h = torch.empty(N, H, device=dev)
        self.pred_v = torch.empty(N, Cn, device=dev)
        self.noise = noise
        self.seed, self.atom_offset = seed, atom_offset
        self.keep_traj = keep_traj
        self.noise_pos = torch.empty(N, 3, device=dev) if noise != 'philox' else None
        self.noise_u = torch.empty(N, Cn, device=dev) if noise != 'philox' else None
        self.log_v0 = torch.empty(N, Cn, device=dev) if keep_traj else None
        self.log_post = torch.empty(N, Cn, device=dev) if keep_traj else None
        self.use_graph = use_graph and not callable(noise)
        self.graph = None
        self._ws = None
        # guidance: dict(cloud=[M,3] float64 device tensor, radius=..., grad_step=..., ratio=0.2, cloud_ptr=None): point-cloud
        # shape guidance of the predicted x0 while t > grad_step ("ShapeMol+g"); the kernel tests t on the device
        self.guidance = guidance
        if keep_traj:
            S = self.num_steps
            per_step = max(1, N) * (12 + 4 + 2 * 4 * Cn)
            if traj_chunk is None:      # ~1 GB per ring half
                traj_chunk = max(1, min(S, (1 << 30) // per_step))
            self.chunk = int(max(1, min(traj_chunk, max(S, 1))))
            spec = (('pos', (N, 3), torch.float32), ('v', (N,), torch.int32), ('v0', (N, Cn), torch.float32), ('vt', (N, Cn), torch.float32))
            self._ring = [{k: torch.empty((self.chunk,) + s, dtype=dt, device=dev) for k, s, dt in spec} for _ in range(2)]
            self.traj = {k: torch.empty((S,) + s, dtype=dt) for k, s, dt in spec}                      # host
            self.traj['pos_cond'] = torch.empty((S, N, 3), dtype=torch.float32, device=dev)             # device (reference :645-646)
            self.traj['v_cond'] = torch.empty((S, N, Cn), dtype=torch.float32, device=dev)
            self._side = torch.cuda.Stream(device=dev)

    def _step_body(self, step):
        e = self.e
        if self.bd.n_atoms == 0:
            return
        # from the second step on the step-independent quantities in the workspace are still valid
        e.forward(self.pos, self.v, self.bd, self.shape, self.t, self.pred_pos, self.pred_h, self.pred_v, reuse_static=step > 0, owner=self)
        if self.guidance is not None:
            gd = self.guidance
            e.guidance(self.bd, self.pred_pos, gd['cloud'], gd['radius'], ratio=gd.get('ratio', 0.2), t_i32=self.t,
                       grad_step=gd['grad_step'], cloud_ptr=gd.get('cloud_ptr'), seed=self.seed, atom_offset=self.atom_offset)
        if self.noise == 'torch':
            self.noise_pos.normal_()     # == torch.randn_like(pos): same generator consumption
            self.noise_u.uniform_()      # == torch.rand_like(logits)
        elif callable(self.noise):
            npos, nu = self.noise(step)
            self.noise_pos.copy_(npos)
            self.noise_u.copy_(nu)
        e.posterior(self.bd, self.pred_pos, self.pred_v, self.t, self.pos, self.v, self.noise_pos, self.noise_u,
                    self.log_v0, self.log_post, seed=self.seed, atom_offset=self.atom_offset)
        e.decrement_t(self.t)

    def _record(self, step):
        tr, ring = self.traj, self._ring[(step // self.chunk) & 1]
        i = step % self.chunk
        tr['pos_cond'][step].copy_(self.pred_pos)     # posterior leaves the predictions intact
        tr['v_cond'][step].copy_(self.pred_v)
        ring['pos'][i].copy_(self.pos)
        ring['v'][i].copy_(self.v)
        ring['v0'][i].copy_(self.log_v0)
        ring['vt'][i].copy_(self.log_post)

    def _drain(self, c, event):
        """Chunk c of the ring -> host, on the side stream (runs beside the main stream's next chunk)."""
        lo = c * self.chunk
        n = min(self.chunk, self.num_steps - lo)
        ring = self._ring[c & 1]
        with torch.cuda.stream(self._side):
            self._side.wait_event(event)
            for k in self.HOST_KEYS:
                self.traj[k][lo:lo + n].copy_(ring[k][:n])      # blocks the host until the bytes have landed
        self._side.synchronize()

    def run(self, progress=None):
        steps = range(self.num_steps)
        if progress is not None:
            steps = progress(steps)
        pending = None            # (chunk index, event recorded after its last step)
        for s in steps:
            if not self.use_graph or s == 0 or self.bd.n_atoms == 0:
                self._step_body(s)        # step 0 runs eagerly (it also warms up the kernels)
            else:
                if self.graph is None:
                    self._capture(s)
                self.graph.replay()
            if self.keep_traj and self.bd.n_atoms:
                self._record(s)           # the trajectory slot depends on the step: outside the graph
                if (s + 1) % self.chunk == 0 or s + 1 == self.num_steps:
                    ev = torch.cuda.Event()
                    ev.record()
                    if pending is not None:
                        self._drain(*pending)     # the ring half the NEXT chunk will overwrite
                    pending = (s // self.chunk, ev)
        if pending is not None:
            self._drain(*pending)
        return self.pos, self.v

    def _capture(self, step):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step_body(step)
        self.graph = g


class HostStepper:
    """One denoising step with HOST buffers on both sides (the end-to-end call bench.py times):
    pinned host (x_t, v_t, t, shape) -> H2D -> network + posterior -> D2H (x_{t-1}, v_{t-1})."""

    def __init__(self, engine, batch_ligand_dev, n_mols, noise='philox', seed=0, use_graph=True):
        m = engine.module
        self.use_graph, self._graph, self._graph_key = use_graph, None, None
        self._ws = None
        self.e = engine
        self.bd = BatchDesc(batch_ligand_dev, n_mols)
        dev = batch_ligand_dev.device
        N, B, Cn, H = self.bd.n_atoms, self.bd.n_mols, m.num_classes, m.hidden_dim
        self.d_pos = torch.empty(N, 3, device=dev)
        self.d_v = torch.empty(N, dtype=torch.int32, device=dev)
        self.d_t = torch.empty(B, dtype=torch.int32, device=dev)
        self.d_shape = torch.empty(B, 32, 3, device=dev)
        self.pred_pos = torch.empty(N, 3, device=dev)
        self.pred_h = torch.empty(N, H, device=dev)
        self.pred_v = torch.empty(N, Cn, device=dev)
        self.h_pos_out = torch.empty(N, 3).pin_memory()
        self.h_v_out = torch.empty(N, dtype=torch.int32).pin_memory()
        self.seed = seed
        self.h2d_bytes = N * 12 + N * 4 + B * 4 + B * 32 * 3 * 4
        self.d2h_bytes = N * 12 + N * 4

    def _enqueue(self, h_pos, h_v, h_t, h_shape):
        self.d_pos.copy_(h_pos, non_blocking=True)
        self.d_v.copy_(h_v, non_blocking=True)
        self.d_t.copy_(h_t, non_blocking=True)
        self.d_shape.copy_(h_shape, non_blocking=True)
        self.e.forward(self.d_pos, self.d_v, self.bd, self.d_shape, self.d_t, self.pred_pos, self.pred_h, self.pred_v, owner=self)
        self.e.posterior(self.bd, self.pred_pos, self.pred_v, self.d_t, self.d_pos, self.d_v, seed=self.seed)
        self.h_pos_out.copy_(self.d_pos, non_blocking=True)
        self.h_v_out.copy_(self.d_v, non_blocking=True)

    def step(self, h_pos, h_v, h_t, h_shape):
        """h_*: pinned host tensors.  Returns pinned host (pos_next, v_next); asynchronous on the
        current stream (synchronise before reading).  The whole step -- the four H2D copies, every kernel and the two D2H
        copies -- is captured once per set of host buffers in a CUDA graph and replayed."""
        key = (h_pos.data_ptr(), h_v.data_ptr(), h_t.data_ptr(), h_shape.data_ptr())
        if self.use_graph and all(t.is_pinned() for t in (h_pos, h_v, h_t, h_shape)):
            if self._graph_key != key:
                self._enqueue(h_pos, h_v, h_t, h_shape)          # eager once: warms up and validates before capture
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(h_pos, h_v, h_t, h_shape)
                self._graph, self._graph_key = g, key
            self._graph.replay()
        else:
            self._enqueue(h_pos, h_v, h_t, h_shape)
        return self.h_pos_out, self.h_v_out
