"""Throughput of the VN-DGCNN shape encoder (BASELINE configs[3]: 1,024-point clouds).
    python tools/bench_encoder.py [--batch 256] [--points 1024]
Prints one JSON line: clouds/s, ms per batch, workspace bytes; the CPU oracle is timed on a 2-cloud sample."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--points', type=int, default=1024)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--cpu', action='store_true')
    args = ap.parse_args()
    from conftest import load_golden
    from test_gpu_encoder import make_ae
    fx = load_golden('encoder.pt')
    ae = make_ae(fx, train=True)
    g = torch.Generator().manual_seed(2021)
    clouds = (3.0 * torch.randn(args.batch, 1, args.points, 3, generator=g)).cuda()
    lat = ae.encoder(clouds)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(lat).all())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        lat = ae.encoder(clouds)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    out = {'metric': 'VN-DGCNN encoder clouds/s', 'value': args.batch / (ms * 1e-3), 'unit': 'clouds/s', 'ms_per_batch': ms,
           'batch': args.batch, 'points': args.points, 'workspace_bytes': int(ae.encoder._ws.numel()),
           'gflop_ref_per_cloud': 35.5 * (args.points / 1024.0) ** 2 if args.points != 1024 else 35.5,
           'ref_tflops': 35.5e9 * args.batch / (ms * 1e-3) / 1e12 if args.points == 1024 else None}
    if args.cpu:
        from oracle import shapemol_oracle as orc
        from test_gpu_encoder import oracle_weights
        w = oracle_weights(fx)
        x = clouds[:2].cpu()
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.encoder_forward(w, x, training=True) if hasattr(orc, 'encoder_forward') else None
        out['cpu_oracle_s_per_cloud'] = (time.perf_counter() - t0) / 2
    print(json.dumps(out))


if __name__ == '__main__':
    main()
