"""GPU parity tests of the VN-DGCNN shape encoder: the CUDA path through the drop-in
PointCloud_AE / C ABI against (1) latents produced by the unmodified reference with the shipped
se_model.pt weights (tests/golden/encoder.pt), (2) the oracle on fresh seeded clouds, (3) the
rotation-equivariance property enc(xQ) = enc(x)Q at a larger size.

Tolerance: the encoder is fp32 end to end; latents within 1e-3 relative (max|d| / max|ref|) whenever the
dynamic kNN graphs agree.  The graph is selected on Gram-matrix distances -|fj|^2 + 2<fi,fj> - |fi|^2 in
R^384, whose fp32 rounding (summation order of the reference's BLAS matmul) decides near-ties at rank k:
on the 512-point fixture the reference's OWN formulas evaluated in fp64 differ from its fp32 output by
2.0e-3 (one swapped neighbour), and the CUDA path lands on the fp64 side.  The test therefore asserts
< 1e-3 against the closer of {reference fp32 fixture, fp64 evaluation of the same formulas} and a hard
< 5e-3 against the fp32 fixture.
"""
import types

import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
REL = 1e-3


def rel_err(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max().clamp_min(1e-12))


def make_ae(fx, train):
    import synth
    from shapemol_b200 import dropin
    dropin.install()
    import models.shape_pointcloud_modelAE as spm
    cfg = types.SimpleNamespace(model_type='PointCloud_AE', encoder='VN_DGCNN', loss_type='signed_distance', latent_dim=32,
                                hidden_dim=128, point_dim=3, layer_num=4, num_k=fx['num_k'])
    ae = spm.PointCloud_AE(cfg)
    sd = {'encoder.' + k: v for k, v in fx['trained'].items()}
    missing, unexpected = ae.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith('generator.') for k in missing)
    bw = synth.synth_state_dict(fx['block_shapes'], 3, skip_non_synth=False)
    for i, b in enumerate(ae.encoder.blocks):
        b.load_state_dict({k[len('blocks.%d.' % i):]: v for k, v in bw.items() if k.startswith('blocks.%d.' % i)})
    ae = ae.to('cuda')
    ae.encoder.conv_pos.train(train)
    ae.encoder.conv_c.train(train)
    return ae


def oracle_weights(fx):
    import synth
    w = dict(fx['trained'])
    w.update(synth.synth_state_dict(fx['block_shapes'], 3, skip_non_synth=False))
    return w


@pytest.mark.parametrize('name', ['small', 'p512'])
@pytest.mark.parametrize('mode', ['train', 'eval'])
def test_encoder_matches_reference_fixture(cuda_lib, name, mode):
    fx = load_golden('encoder.pt')
    ae = make_ae(fx, mode == 'train')
    lat = ae.encoder(fx['%s_clouds' % name].cuda())
    torch.cuda.synchronize()
    ref = fx['%s_%s_latent' % (name, mode)]
    assert lat.shape == ref.shape
    e = rel_err(lat, ref)
    from oracle import shapemol_oracle as orc
    w64 = {k: (v.double() if v.is_floating_point() else v) for k, v in oracle_weights(fx).items()}
    with torch.no_grad():
        ref64 = orc.encoder_forward(w64, fx['%s_clouds' % name].double(), k=fx['num_k'], training=(mode == 'train'))
    e64 = rel_err(lat, ref64)
    print('encoder %s/%s: rel err %.2e vs reference fp32 fixture, %.2e vs fp64 evaluation' % (name, mode, e, e64))
    assert min(e, e64) < REL and e < 5e-3


def test_encoder_vs_oracle_seeded_and_bn_side_effects(cuda_lib):
    from oracle import shapemol_oracle as orc
    fx = load_golden('encoder.pt')
    g = torch.Generator().manual_seed(77)
    clouds = torch.randn(3, 1, 200, 3, generator=g) * torch.tensor([2.5, 1.5, 1.0])   # P not a multiple of 32 / 128
    clouds = clouds - clouds.mean(2, keepdim=True)
    w = oracle_weights(fx)
    for train in (True, False):
        ae = make_ae(fx, train)
        bn = ae.encoder.conv_pos.batchnorm.bn
        rm0, nbt0 = bn.running_mean.clone(), int(bn.num_batches_tracked)
        lat = ae.encoder(clouds.cuda())
        with torch.no_grad():
            ref = orc.encoder_forward(w, clouds, k=fx['num_k'], training=train)
        w64 = {k: (v.double() if v.is_floating_point() else v) for k, v in w.items()}
        with torch.no_grad():
            ref64 = orc.encoder_forward(w64, clouds.double(), k=fx['num_k'], training=train)
        e, e64 = rel_err(lat, ref), rel_err(lat, ref64)
        print('encoder vs oracle (train=%s): rel err %.2e (fp32 oracle) %.2e (fp64 oracle)' % (train, e, e64))
        assert min(e, e64) < REL and e < 5e-3
        if train:   # nn.BatchNorm side effects of a train-mode pass
            assert int(bn.num_batches_tracked) == nbt0 + 1 and not torch.equal(bn.running_mean, rm0)
        else:
            assert int(bn.num_batches_tracked) == nbt0 and torch.equal(bn.running_mean, rm0)
        # the unregistered blocks always use (and update) batch statistics
        assert int(ae.encoder.blocks[0].batchnorm.bn.num_batches_tracked) == 1


def test_encoder_rotation_equivariance_at_1024_points(cuda_lib):
    """enc(xQ) = enc(x)Q: a size-independent property checked at the BASELINE config-4
    cloud size (1,024 points; batch bounded so the test stays in seconds)."""
    fx = load_golden('encoder.pt')
    ae = make_ae(fx, True)
    g = torch.Generator().manual_seed(5)
    clouds = torch.randn(8, 1, 1024, 3, generator=g) * torch.tensor([3.0, 2.0, 1.5])
    clouds = clouds - clouds.mean(2, keepdim=True)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    lat = ae.encoder(clouds.cuda()).cpu()
    lat_rot = ae.encoder((clouds @ q).cuda()).cpu()
    assert torch.isfinite(lat).all()
    e = rel_err(lat_rot, lat @ q)
    print('encoder equivariance at P=1024: rel err %.2e' % e)
    assert e < REL


def test_encoder_argument_errors(cuda_lib):
    from shapemol_b200 import _lib
    fx = load_golden('encoder.pt')
    ae = make_ae(fx, True)
    with pytest.raises(_lib.SmbError):
        ae.encoder(torch.zeros(1, 1, 8, 3, device='cuda'))        # fewer points than num_k
    with pytest.raises(_lib.SmbError):
        ae.encoder(torch.zeros(1, 1, 64, 3))                       # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ae.encoder(torch.zeros(1, 64, 3, device='cuda'))
    assert ae.encoder(torch.zeros(0, 1, 64, 3, device='cuda')).shape == (0, 32, 3)


def test_frontend_many_shapes_end_to_end(cuda_lib):
    """molecules -> mesh-free surface clouds -> centred / bounded / batched -> CUDA encoder, through
    shapemol_b200.shape_frontend.get_pointAE_shape_emb (reference utils/shape.py:240-284).  The latents of every batch are
    checked against the oracle encoder evaluated on exactly the clouds the front end produced."""
    import synth
    from oracle import shapemol_oracle as orc
    from shapemol_b200 import shape_frontend as sf
    fx = load_golden('encoder.pt')
    ae = make_ae(fx, True)
    g = torch.Generator(device='cuda').manual_seed(9)
    sizes = [13, 27, 9, 21, 18]
    xyz = synth.molecule_like_positions(sizes, 31)
    mols, o = [], 0
    for n in sizes:
        mols.append(dict(coords=xyz[o:o + n].cuda(), atomic_numbers=[6, 7, 8][o % 3:o % 3 + 1] * n))
        o += n
    zs, bounds, clouds, centers = sf.get_pointAE_shape_emb(mols, ae, 256, batch_size=2, generator=g)
    assert zs.shape == (5, 32, 3) and bounds.shape == (5, 3, 2) and centers.shape == (5, 3)
    assert [tuple(c.shape) for c in clouds] == [(2, 1, 256, 3), (2, 1, 256, 3), (1, 1, 256, 3)]
    assert all(float(c.mean(dim=2).abs().max()) < 1e-4 for c in clouds)            # centred
    assert bool((bounds[:, :, 0] < 0).all() and (bounds[:, :, 1] > 0).all())       # the box contains the centre
    w = oracle_weights(fx)
    w64 = {k: (v.double() if v.is_floating_point() else v) for k, v in w.items()}
    o = 0
    for c in clouds:
        with torch.no_grad():
            ref = orc.encoder_forward(w, c, k=fx['num_k'], training=True)
            ref64 = orc.encoder_forward(w64, c.double(), k=fx['num_k'], training=True)
        e, e64 = rel_err(zs[o:o + c.shape[0]], ref), rel_err(zs[o:o + c.shape[0]], ref64)
        print('front end batch of %d: rel err %.2e (fp32 oracle) %.2e (fp64 oracle)' % (c.shape[0], e, e64))
        # surface clouds sampled from overlapping spheres contain near-coincident points: the fp32 and the fp64 evaluation of
        # the oracle itself disagree on such neighbours (up to ~1e-2 here), so the bar is "as close to either oracle as they
        # are to each other" rather than 1e-3
        assert min(e, e64) < max(REL, 1.2 * rel_err(ref, ref64)) and e < 2e-2
        o += c.shape[0]
