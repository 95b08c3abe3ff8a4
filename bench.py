#!/usr/bin/env python
"""bench.py -- throughput of the reverse-diffusion denoising step (BASELINE.json metric).

A "step" is one denoising step (network evaluation + posterior update) over one batch.
Workload at N=1: BASELINE configs[1] -- 100 condition shapes x 50 molecules, atom counts drawn from the
MOSES2 size prior (9..27 heavy atoms, mean 21.4), k=32, hidden 128, 8 layers, train-mode BatchNorm as
scripts/sample_diffusion.py runs it.  N>1 (torchrun): every rank runs the same-sized batch of its own
molecules (weak scaling, no per-step communication) and the final states are gathered once with NCCL.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python bench.py --impl reference ...      # the CPU arm (oracle port of the reference's PyTorch path)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic work of the reference formulation, per edge of one edge-MLP + logits:
# Linear(308->128) + Linear(128->128) + <q,k>
F_REF_EDGE_K = 2 * (308 * 128 + 128 * 128) + 2 * 128
F_MIN_EDGE_K = 2 * (20 * 128 + 128 * 128) + 2 * 128       # with the first Linear factored to node level
KERNELS_PER_STEP = 4 + 8 * 9 + 1 + 1 + 1                   # prep, embed, knn, gate | 8 x (3 node, 4 edge, bn, apply) | head | posterior | t--
#                                                            (the tile list is built once: steps after the first reuse it)


def ref_like_config(k=32):
    import types
    return types.SimpleNamespace(
        denoise_type='diffusion', model_mean_type='C0', topo_emb_type=None, gt_noise_type='origin',
        schedule_pos=dict(beta_schedule='sigmoid', beta_start=1.e-7, beta_end=0.01, s=6),
        schedule_v=dict(beta_schedule='cosine', s=0.01), num_diffusion_timesteps=1000, loss_v_weight=100.0,
        v_mode='uniform', v_net_type='mlp', loss_pos_type='mse', sample_time_method='symmetric',
        loss_weight_type='noise_level', loss_pos_min_weight=0, loss_pos_max_weight=10, time_emb_dim=8,
        center_pos_mode='none', atom_enc_mode='add_aromatic', model_type='uni_o2', num_blocks=1, num_layers=8,
        hidden_dim=128, n_heads=16, edge_feat_dim=0, num_r_gaussian=20, knn=k, num_node_types=8, act_fn='relu',
        norm=True, cutoff_mode='knn', ew_net_type='global', r_feat_mode='sparse', num_x2h=1, num_h2x=1, r_max=10.0,
        x2h_out_fc=False, sync_twoup=False, shape_dim=32, shape_latent_dim=32, shape_type='pointAE_shape',
        cond_mask_prob=0.0)


def make_workload(n_shapes, per_shape, seed, fixed_atoms=0):
    """Synthetic MOSES-shaped batch: ragged atom counts from the size prior, N(0,1) initial positions,
    uniform initial types, N(0, 0.07^2) shape latents (one per shape, repeated per molecule)."""
    with open(os.path.join(ROOT, 'data', 'moses_atom_count_hist.json')) as f:
        hist = json.load(f)['counts']
    ns = torch.tensor([int(k) for k in hist], dtype=torch.long)
    w = torch.tensor([float(v) for v in hist.values()])
    g = torch.Generator().manual_seed(seed)
    B = n_shapes * per_shape
    sizes = ns[torch.multinomial(w, B, replacement=True, generator=g)]
    if fixed_atoms > 0:      # BASELINE configs[2]: every molecule has exactly this many atoms
        sizes = torch.full((B,), int(fixed_atoms), dtype=torch.long)
    N = int(sizes.sum())
    batch = torch.repeat_interleave(torch.arange(B), sizes)
    pos = torch.randn(N, 3, generator=g)
    v = torch.randint(0, 15, (N,), generator=g)
    shape = (0.07 * torch.randn(n_shapes, 32, 3, generator=g)).repeat_interleave(per_shape, dim=0)
    return sizes, batch, pos, v, shape


def build_model(k, precision):
    from shapemol_b200 import dropin
    dropin.install()
    import models.molopt_score_model as msm
    torch.manual_seed(2021)
    m = msm.ScorePosNet3D(ref_like_config(k), ligand_atom_feature_dim=15)
    m.smb_precision = precision
    return m


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self.stop_flag, self.index = [], False, index

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
            'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            for name, col in (('hw_slowdown', 3), ('hw_thermal_slowdown', 4), ('sw_thermal_slowdown', 5), ('sw_power_cap', 6)):
                if len(r) > col and r[col].lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_reference_arm(k, n_mols, steps, warmup, seed=2021):
    """The reference's PyTorch CPU path as restated by the oracle port (the Python reference cannot
    travel to the GPU box).  Returns (seconds per step, threads, description)."""
    from oracle import shapemol_oracle as orc
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = build_model(k, 'bf16x3')
    sd = {kk: vv.detach().clone() for kk, vv in m.state_dict().items()}
    cfg = dict(orc.DEFAULT_CFG, knn=k)
    tabs = orc.schedule_tables(1000, cfg['schedule_pos'], cfg['schedule_v'])
    sizes, batch, pos, v, shape = make_workload(max(1, n_mols // 50), min(50, n_mols), seed)
    mol_ptr = orc.mol_ptr_from_sizes(sizes.tolist())
    B = sizes.numel()
    g = torch.Generator().manual_seed(1)
    times = []
    with torch.no_grad():
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            t = torch.full((B,), 999 - s, dtype=torch.long)
            x0, _, logits = orc.forward(sd, cfg, pos, v, mol_ptr, shape, t, training=True)
            eps, u = torch.randn(pos.shape, generator=g), torch.rand(logits.shape, generator=g)
            pos, v, _, _ = orc.posterior_step(tabs, x0, logits, pos, v, t[batch], eps, u)
            if s >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), threads, B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['bf16x3', 'bf16'])
    ap.add_argument('--no-parity-mode', action='store_true', help='skip the secondary bf16x3 (fp32-parity) measurement')
    ap.add_argument('--shapes', type=int, default=100)
    ap.add_argument('--per-shape', type=int, default=50)
    ap.add_argument('--k', type=int, default=32)
    ap.add_argument('--fixed-atoms', type=int, default=0, help='all molecules of this size (configs[2]: 27) instead of the MOSES size prior')
    ap.add_argument('--cpu-mols', type=int, default=50)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--prof-kernel', default='edge_k')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    workload = '%d shapes x %d molecules, %s, k=%d, hidden 128, 8 layers, T=1000, train-mode BN' % (
        args.shapes, args.per_shape, ('%d atoms each' % args.fixed_atoms) if args.fixed_atoms else 'MOSES size prior (9..27 atoms)', args.k)

    if args.impl == 'reference':
        if rank != 0:
            return
        sec, threads, bc = cpu_reference_arm(args.k, args.cpu_mols, args.steps, args.warmup)
        val = bc / (1000.0 * sec)
        print(json.dumps({
            'impl': 'reference', 'metric': 'molecules/sec (1000-step sampling)', 'value': val, 'unit': 'molecules/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload, 'sample': '%d molecules per step on the host CPU' % bc},
            'mol_steps_per_s': bc / sec,
            'cpu_baseline': {'value': val, 'unit': 'molecules/s', 'cores': threads, 'kind': 'port',
                             'sample': '%d molecules x %d timed steps, oracle port of the reference PyTorch path' % (bc, args.steps)},
            'e2e': {'value': val, 'unit': 'molecules/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    from shapemol_b200.engine import Sampler, HostStepper

    model = build_model(args.k, args.precision).to(dev).train()
    sizes, batch, pos, v, shape = make_workload(args.shapes, args.per_shape, 2021 + rank, args.fixed_atoms)
    B, N = sizes.numel(), int(sizes.sum())
    E = int((sizes * torch.clamp(sizes - 1, max=args.k)).sum())
    eng = model._engine()
    atom_offset = rank * N
    sampler = Sampler(eng, pos.to(dev), v.to(dev), batch.to(dev), shape.to(dev), num_steps=1000, noise='philox', seed=2021,
                      atom_offset=atom_offset, keep_traj=False, use_graph=True, n_mols=B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: step 0 eager, then capture + replay ----
    sampler._step_body(0)
    sampler._capture(1)
    for _ in range(args.warmup):
        sampler.graph.replay()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        sampler.graph.replay()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms)
    assert bool(torch.isfinite(sampler.pos).all()), 'non-finite coordinates'

    # ---- one final gather of the results (the only collective of the path) ----
    gather_ms = None
    if world > 1:
        packed = torch.cat([sampler.pos.flatten(), sampler.v.float()])
        cnt = torch.tensor([packed.numel()], device=dev)
        cnts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        mx = int(max(int(c) for c in cnts))
        buf = torch.zeros(mx, device=dev)
        buf[:packed.numel()] = packed
        out = torch.empty(world * mx, device=dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        dist.all_gather_into_tensor(out, buf)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)

    # ---- end to end through host buffers (H2D of the step's inputs, D2H of its result, every step) ----
    hs = HostStepper(eng, batch.to(dev), B, seed=2021)
    h_pos, h_v = pos.clone().pin_memory(), v.to(torch.int32).pin_memory()
    h_t, h_shape = torch.full((B,), 500, dtype=torch.int32).pin_memory(), shape.clone().pin_memory()
    for _ in range(args.warmup):
        hs.step(h_pos, h_v, h_t, h_shape)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        hs.step(h_pos, h_v, h_t, h_shape)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / args.steps
    clocks.stop_flag = True          # sampled through both timed regions (device-resident loop and host-buffer loop)
    clocks.join(timeout=3)
    if world > 1:
        tms = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = float(tms)

    # ---- live per-kernel timing of the dominant kernel (roofline line) ----
    roof = None
    if rank == 0:
        n_launch = {'edge_k': 16, 'edge_v': 8, 'edge_xv': 8, 'node_pre': 16, 'node_out': 8}.get(args.prof_kernel, 1)
        events = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n_launch)]
        for ev in events:
            ev.record()
        durs = []
        for rep in range(4):
            eng.forward(sampler.pos, sampler.v, sampler.bd, sampler.shape, sampler.t, sampler.pred_pos, sampler.pred_h,
                        sampler.pred_v, prof=(args.prof_kernel, events))
            torch.cuda.synchronize()
            if rep:
                durs += [events[2 * i].elapsed_time(events[2 * i + 1]) for i in range(n_launch)]
        kms = sum(durs) / len(durs)
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback (B200_PROFILING.md sustained)'
        hbm = peaks.get('hbm_gbs', 6650.0)
        if args.prof_kernel == 'edge_k':
            flops, fmin = E * F_REF_EDGE_K, E * F_MIN_EDGE_K
            # reads: A,B projections (bf16 images: 2*128*2 per atom; fp32 in the bf16x3 path), Q (512), x, nbr, e_w;
            # writes alpha*e_w (64 B per edge)
            proj = 2 * 256 if args.precision == 'bf16' else 2 * 512
            bytes_alg = N * (proj + 512 + 12 + 2 * 4 * (args.k + 1)) + E * 64
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full, profiles/r1_v4_ncu_full_summary.txt);
            # only valid for the default workload in bf16 mode
            traffic = 240.0e6 if (args.precision == 'bf16' and args.shapes == 100 and args.per_shape == 50 and args.k == 32 and not args.fixed_atoms) else None
            roof = {'kernel': ('edge_ws_kernel<ROLE_K>' if args.precision == 'bf16' else 'edge_kernel<ROLE_K>') + ' (edge MLP + attention logits + per-destination softmax)',
                    'bound': 'tensor', 'achieved': flops / (kms * 1e-3) / 1e12, 'peak': peak_tf, 'unit': 'TFLOP/s',
                    'frac': flops / (kms * 1e-3) / 1e12 / peak_tf, 'traffic': traffic, 'peak_source': peak_src,
                    'ms_per_launch': kms, 'launches_per_step': 16, 'share_of_step': 16 * kms / ms,
                    'flops_per_launch_ref': flops, 'flops_per_launch_factored': fmin,
                    'achieved_factored_tflops': fmin / (kms * 1e-3) / 1e12,
                    'hbm_bytes_per_launch': bytes_alg, 'hbm_gbs': bytes_alg / (kms * 1e-3) / 1e9, 'hbm_frac': bytes_alg / (kms * 1e-3) / 1e9 / hbm}
        else:
            roof = {'kernel': args.prof_kernel, 'ms_per_launch': kms, 'launches_per_step': n_launch, 'share_of_step': n_launch * kms / ms}

    # ---- secondary: the fp32-parity arithmetic mode (split-bf16 products) on the same workload ----
    parity_mode = None
    if rank == 0 and world == 1 and args.precision == 'bf16' and not args.no_parity_mode:
        m3 = build_model(args.k, 'bf16x3').to(dev).train()
        s3 = Sampler(m3._engine(), pos.to(dev), v.to(dev), batch.to(dev), shape.to(dev), num_steps=1000, noise='philox', seed=2021,
                     atom_offset=atom_offset, keep_traj=False, use_graph=True, n_mols=B)
        s3._step_body(0)
        s3._capture(1)
        for _ in range(3):
            s3.graph.replay()
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(5):
            s3.graph.replay()
        p1.record()
        torch.cuda.synchronize()
        ms3 = p0.elapsed_time(p1) / 5
        parity_mode = {'dtype': 'bf16x3 (split-bf16 products, fp32 accumulate): x0 / logits within 1e-3 of the fp32 reference',
                       'ms_per_step': ms3, 'mol_steps_per_s': B / (ms3 * 1e-3), 'steps': 5}
        del s3, m3

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, threads, bc = cpu_reference_arm(args.k, args.cpu_mols, 3, 1)
        cpu = {'value': bc / (1000.0 * sec), 'unit': 'molecules/s', 'cores': threads, 'kind': 'port',
               'sample': '%d molecules x 3 timed steps (1 warm-up) of the same workload, oracle port, %.2f s/step' % (bc, sec),
               'mol_steps_per_s': bc / sec}

    if rank == 0:
        total_mols = world * B
        val = total_mols / (1000.0 * ms * 1e-3)
        line = {
            'metric': 'molecules/sec (1000-step sampling)', 'value': val, 'unit': 'molecules/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16x3 (split-bf16 tensor-core products, fp32 accumulate; fp32-parity mode)'
            if args.precision == 'bf16x3' else 'bf16 (tcgen05 MLP contractions, fp32 accumulate; everything else fp32)',
            'data': 'synthetic',
            'config': {'workload': workload, 'molecules_per_gpu': B, 'atoms_per_gpu': N, 'edges_per_gpu': E,
                       'l2': 'per-step working set %.0f MB > 126 MB L2 (no flush needed)' % (N * 7.2e3 / 1e6),
                       'noise': 'in-kernel Philox', 'graph': 'CUDA graph replay per step'},
            'mol_steps_per_s': total_mols / (ms * 1e-3), 'us_per_step': ms * 1e3,
            'mol_steps_per_s_per_gpu': B / (ms * 1e-3),
            'clocks': clocks.summary(),
            'e2e': {'value': total_mols / (1000.0 * e2e_ms * 1e-3), 'unit': 'molecules/s', 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': hs.h2d_bytes, 'd2h_bytes_per_step': hs.d2h_bytes},
            'gpu_launches': KERNELS_PER_STEP * args.steps,
            'roofline': roof, 'cpu_baseline': cpu, 'fp32_parity_mode': parity_mode, 'final_gather_ms': gather_ms,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
