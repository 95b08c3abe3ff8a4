import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/tests/golden')
import torch
from conftest import load_golden
from test_gpu_parity import build_model, batch_of
fx = load_golden('forward_tiny_train.pt')
m = build_model(fx, 'bf16')
out = m(fx['pos'].cuda(), fx['v'].cuda(), batch_of(fx['sizes']), fx['shape'].cuda(), time_step=fx['t'].cuda())
torch.cuda.synchronize()
print('ok', float(out['pred_ligand_pos'].abs().max()))
