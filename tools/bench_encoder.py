"""Throughput of the VN-DGCNN shape encoder (BASELINE configs[3]: 4,096 clouds x 1,024 points).

    python tools/bench_encoder.py [--clouds 4096] [--chunk 256] [--points 1024]

`measure()` is also the `encoder` leg of bench.py.  The batch is processed in chunks of `chunk` clouds (the per-cloud
workspace is ~15 MB); the DGCNN blocks' BatchNorm statistics (always batch statistics) are therefore per chunk --
the reference run with batch = chunk.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)

# algorithmic work per cloud (SURVEY 8d, P = 1024, kk = 20, Hs = 128, 4 blocks)
GFLOP_REF, GFLOP_FACTORED, MB_MIN = 35.5, 4.9, 19.0


def measure(dev, clouds=4096, chunk=256, points=1024, reps=1, cpu_clouds=0):
    from conftest import load_golden
    from test_gpu_encoder import make_ae
    fx = load_golden('encoder.pt')
    ae = make_ae(fx, train=True)
    g = torch.Generator().manual_seed(2021)
    # unit-scale molecular surfaces: points on spheres of radius 2..4 A around a few centres, synthetic
    x = torch.randn(chunk, 1, points, 3, generator=g)
    x = (x / x.norm(dim=-1, keepdim=True)) * (2.0 + 2.0 * torch.rand(chunk, 1, points, 1, generator=g))
    x = x.to(dev)
    n_chunks = max(1, clouds // chunk)
    lat = ae.encoder(x)                      # warm-up (allocates the workspace)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(lat).all())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for c in range(n_chunks):
            lat = ae.encoder(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    total = n_chunks * chunk
    scale = (points / 1024.0)
    peak_tf, hbm = 1400.0, 6650.0
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            pk = json.load(f)
        peak_tf, hbm = pk.get('bf16_tflops_sustained', peak_tf), pk.get('hbm_gbs', hbm)
    except Exception:
        pass
    out = {'metric': 'VN-DGCNN encoder clouds/s (BASELINE configs[3])', 'value': total / (ms * 1e-3), 'unit': 'clouds/s',
           'clouds': total, 'points': points, 'chunk': chunk, 'ms_total': ms, 'ms_per_chunk': ms / n_chunks,
           'workspace_bytes': int(ae.encoder._ws.numel()), 'dtype': 'f32',
           'note': 'BatchNorm batch statistics per chunk of %d clouds' % chunk,
           'roofline': {'bound': 'hbm', 'unit': 'GB/s', 'peak': hbm,
                        'achieved': MB_MIN * scale * 1e6 * total / (ms * 1e-3) / 1e9,
                        'frac': MB_MIN * scale * 1e6 * total / (ms * 1e-3) / 1e9 / hbm,
                        'traffic': None, 'bytes_per_cloud_algorithmic': MB_MIN * scale * 1e6,
                        'tflops_factored': GFLOP_FACTORED * scale * scale * 1e9 * total / (ms * 1e-3) / 1e12,
                        'tflops_ref_formulation': GFLOP_REF * scale * scale * 1e9 * total / (ms * 1e-3) / 1e12,
                        'frac_of_bf16_peak_factored': GFLOP_FACTORED * scale * scale * 1e9 * total / (ms * 1e-3) / 1e12 / peak_tf}}
    if cpu_clouds > 0:
        sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
        import ref_loader
        if ref_loader.available():
            _, spm, ED = ref_loader.load()
            ck = torch.load(os.path.join(ref_loader.REF_ROOT, 'trained_models', 'se_model.pt'), map_location='cpu', weights_only=False)
            ref = spm.PointCloud_AE(ck['config'].model)
            ref.load_state_dict(ck['model'], strict=True)
            xc = x[:cpu_clouds].cpu()
            torch.set_num_threads(os.cpu_count() or 1)
            t0 = time.perf_counter()
            with torch.no_grad():
                ref.encoder(xc)
            out['cpu_baseline'] = {'value': cpu_clouds / (time.perf_counter() - t0), 'unit': 'clouds/s', 'cores': os.cpu_count() or 1,
                                   'kind': 'reference', 'sample': '%d clouds x %d points, unmodified reference VN_DGCNN_Encoder' % (cpu_clouds, points)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--clouds', type=int, default=4096)
    ap.add_argument('--chunk', type=int, default=256)
    ap.add_argument('--points', type=int, default=1024)
    ap.add_argument('--reps', type=int, default=1)
    ap.add_argument('--cpu-clouds', type=int, default=0)
    args = ap.parse_args()
    print(json.dumps(measure(torch.device('cuda', 0), args.clouds, args.chunk, args.points, args.reps, args.cpu_clouds)))


if __name__ == '__main__':
    main()
