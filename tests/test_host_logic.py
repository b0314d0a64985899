"""CPU: host-side logic of the drop-in surface — blueprint registry, state_dict compatibility with the
reference, conv geometry, batch sharding."""
import copy

import pytest
import torch

import open_genie_b200 as og
from open_genie_b200 import ops
from open_genie_b200.ddp import shard_batch
from open_genie_b200.module import get_module, parse_blueprint
from oracle import fixtures as fx


def test_parse_blueprint_expands_and_does_not_mutate():
    bp = (('video-residual', {'n_rep': 3, 'in_channels': 64}),
          ('adaptive_group_norm', {'dim_cond': 6, 'num_groups': 8, 'num_channels': 64, 'has_ext': True}),
          'silu')
    before = copy.deepcopy(bp)
    layers, ext = parse_blueprint(bp)
    assert len(layers) == 5 and ext == [False, False, False, True, False]
    assert bp == before                      # the reference pops n_rep/has_ext from the caller's dicts; we do not
    layers2, _ = parse_blueprint(bp)
    assert len(layers2) == 5


def test_registry_names():
    for name in ('space_attn', 'time_attn', 'space-time_attn', 'video-residual', 'causal-conv3d',
                 'depth2spacetime_upsample', 'spacetime_downsample', 'group_norm', 'adaptive_group_norm', 'silu',
                 'blur_pool'):
        assert isinstance(get_module(name), type)
    with pytest.raises(ValueError, match='Unknown module name'):
        get_module('nope')
    with pytest.raises(NotImplementedError):
        get_module('causal-conv3d-transpose')


def test_state_dict_matches_the_reference(golden):
    g = golden('tokenizer_mini.pt')
    tok = og.VideoTokenizer(fx.MINI_ENC, fx.MINI_DEC, d_codebook=fx.MINI_D_CODEBOOK, gan_loss_weight=0,
                            perc_loss_weight=0)
    ref_keys = set(g['grads']['norm'])                   # every parameter name of the reference module
    assert {k for k, _ in tok.named_parameters()} == ref_keys
    assert 'quant.bit_mask' in tok.state_dict()
    for k, v in g['grads']['full'].items():
        assert tuple(dict(tok.named_parameters())[k].shape) == tuple(v.shape), k
    full = og.VideoTokenizer(og.MAGVIT2_ENC_DESC, og.MAGVIT2_DEC_DESC, gan_loss_weight=0, perc_loss_weight=0)
    assert sum(p.numel() for p in full.parameters()) == 375_554_837
    assert len(full.enc_layers) == 27 and len(full.dec_layers) == 31      # layers after n_rep expansion
    w = full.enc_layers[1].main[2].weight
    assert w.shape == (128, 128, 3, 3, 3) and w.permute(0, 2, 3, 4, 1).is_contiguous()   # [Cout][tap][Cin] bytes
    # a reference-format (contiguous NCDHW-order) state_dict loads, values land in the right logical slots
    sd = {k: torch.randn_like(v).contiguous() for k, v in tok.state_dict().items() if v.dtype.is_floating_point}
    tok.load_state_dict(sd, strict=False)
    k = 'enc_layers.1.main.2.weight'
    assert torch.equal(tok.state_dict()[k], sd[k])


def test_tokenizer_builds_the_auxiliary_losses_like_the_reference():
    """genie/tokenizer.py:288-299: GANLoss(FrameDiscriminator(**disc_kwargs)) and PerceptualLoss(vgg16) when their
    weights are > 0 (the constructor defaults), nn.Identity otherwise; 'frames' and 'video' critics."""
    import torch.nn as nn
    from open_genie_b200.module.loss import GANLoss, PerceptualLoss
    tok = og.VideoTokenizer(fx.MINI_ENC, fx.MINI_DEC, d_codebook=6, disc_kwargs={'inp_size': 32})
    assert isinstance(tok.gan_crit, GANLoss) and isinstance(tok.perc_crit, PerceptualLoss)
    keys = set(tok.state_dict())
    assert {'gan_crit.disc.proj_in.weight', 'gan_crit.disc.core.1.0.main.6.go_up.1.weight',
            'gan_crit.disc.to_logits.3.weight', 'perc_crit.percept_model.features.0.weight',
            'perc_crit.percept_model.features.28.bias'} <= keys
    assert tok.state_dict()['gan_crit.disc.proj_in.weight'].shape == (64, 3, 3, 3)        # nn.Conv2d layout
    assert not any(p.requires_grad for p in tok.perc_crit.parameters())                   # frozen extractor (loss.py:49-51)
    tok0 = og.VideoTokenizer(fx.MINI_ENC, fx.MINI_DEC, d_codebook=6, gan_loss_weight=0, perc_loss_weight=0)
    assert isinstance(tok0.gan_crit, nn.Identity) and isinstance(tok0.perc_crit, nn.Identity)
    from open_genie_b200.module.discriminator import VideoDiscriminator
    tokv = og.VideoTokenizer(fx.MINI_ENC, fx.MINI_DEC, d_codebook=6, gan_discriminate='video', perc_loss_weight=0,
                             disc_kwargs={'inp_size': (8, 32)})
    assert isinstance(tokv.gan_crit.disc, VideoDiscriminator) and tokv.gan_crit.disc.proj_in.weight.shape == (64, 3, 3, 3, 3)
    with pytest.raises(NotImplementedError, match='use_attn'):
        VideoDiscriminator(inp_size=(8, 32), use_attn=True)


def test_conv_geometry():
    g = ops.ConvGeom(128, 128, (3, 3, 3), causal=True)
    assert (g.pt, g.ph, g.pw, g.direct) == (2, 1, 1, True)                    # video.py:155: (kt-1) + (1 - stride)
    g = ops.ConvGeom(128, 128, (3, 3, 3), causal=False)
    assert (g.pt, g.direct) == (1, True)
    g = ops.ConvGeom(128, 128, (3, 3, 3), stride=(2, 2, 2))
    assert (g.pt, g.direct, g.strided_implicit, g.out_dims(16, 32, 32)) == (1, False, True, (8, 16, 16))
    g = ops.ConvGeom(3, 128, (3, 3, 3))            # narrow Cin: implicit GEMM on a channel-padded input
    assert g.direct and g.padded and g.cin_pad == 64 and g.kpad == 27 * 64 and g.k_main == 81
    g = ops.ConvGeom(18, 512, (3, 3, 3))
    assert g.direct and g.cin_pad == 64 and g.k_main == 18 * 27
    g = ops.ConvGeom(128, 128, (3, 3, 3), stride=(1, 2, 2))
    assert g.out_dims(16, 64, 64) == (16, 32, 32)


def test_shard_batch_partitions_clips():
    for world in (1, 2, 4, 8):
        spans = [shard_batch(32, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 32
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert [shard_batch(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_zero_arena_lifetime_contract():
    """ops.ZeroArena: zeroed views of segment memory; the step after the first consolidates into one buffer; a view is
    recycled (and re-zeroed) by the first allocation after mark_step()."""
    import torch
    from open_genie_b200 import ops
    a = ops.ZeroArena()
    a.SEGMENT = 4096                                                          # small segments: force growth
    x = a.zeros((3, 4), torch.float32, 'cpu')
    big = a.zeros((4096,), torch.float32, 'cpu')                              # does not fit: a second segment
    assert x.sum() == 0 and big.sum() == 0 and len(a.segs[torch.device('cpu')]) == 2
    assert a.locate(x) == (0, 0) and a.locate(big) == (1, 0)
    x += 1
    big += 2
    a.mark_step()
    y = a.zeros((3, 4), torch.float32, 'cpu')                                 # boundary: consolidated, everything zero
    z = a.zeros((4096,), torch.float64, 'cpu')
    assert len(a.segs[torch.device('cpu')]) == 1 and y.sum() == 0 and z.sum() == 0
    assert y.untyped_storage().data_ptr() == z.untyped_storage().data_ptr()  # one buffer
    y += 1
    z += 2
    a.mark_step()
    w = a.zeros((3, 4), torch.float32, 'cpu')
    assert w.data_ptr() == y.data_ptr() and w.sum() == 0                      # recycled and cleared
    assert a.bytes_in_use() == 256


def test_zero_arena_scoping_contract():
    """a step's accumulators must never be handed to forward-only / no_grad work, and a private
    (graph-owned) scope must neither grow nor be shared with the default scope."""
    import torch
    from open_genie_b200 import ops
    ops.enable_zero_arena(True)
    try:
        a = ops.current_arena()
        with torch.no_grad():
            z = ops._zeros((16,), torch.float32, 'cpu')
        assert a.locate(z) is None                                  # inference / validation forward: plain torch.zeros
        g = ops._zeros((16,), torch.float32, 'cpu')                 # grad mode on (training forward): arena
        assert a.locate(g) is not None
        g += 1.0
        ops.mark_step()
        with torch.no_grad():
            z2 = ops._zeros((16,), torch.float32, 'cpu')
        assert a.locate(z2) is None and float(z2.sum()) == 0.0      # even right after a step boundary
        g2 = ops._zeros((16,), torch.float32, 'cpu')
        assert float(g2.sum()) == 0.0                               # recycled range was cleared
        # a private scope (what GraphedTrainStep owns) is separate memory and refuses to grow once frozen
        scope = ops.StepScope(ops.ZeroArena())
        with ops.step_scope(scope):
            p = ops._zeros((1024,), torch.float32, 'cpu')
            assert scope.arena.locate(p) is not None and a.locate(p) is None
            ws = ops._workspace(torch.device('cpu'), 1 << 20)
        scope.freeze()
        assert ops.current_arena() is a                              # default scope restored
        with ops.step_scope(scope):
            ops.mark_step()
            assert scope.arena.locate(ops._zeros((1024,), torch.float32, 'cpu')) is not None     # same range again
            assert ops._workspace(torch.device('cpu'), 1 << 20).data_ptr() == ws.data_ptr()
            import pytest
            with pytest.raises(RuntimeError, match='too small'):
                ops._zeros((1 << 28,), torch.float32, 'cpu')
            with pytest.raises(RuntimeError, match='would have to grow'):
                ops._workspace(torch.device('cpu'), 1 << 29)
    finally:
        ops.enable_zero_arena(False)


def test_conv_modules_survive_deepcopy():
    """a deep copy must register its own weights for the fused optimizer and re-link its fused shortcut."""
    import copy
    from open_genie_b200.module.video import CONV_REGISTRY, VideoResidualBlock
    m = VideoResidualBlock(64, 128)
    c = copy.deepcopy(m)
    for mod in (c.main[2], c.main[6], c.res[1]):
        assert CONV_REGISTRY.get(id(mod.weight)) is mod
    assert c.main[6]._extra is c.res[1] and c.res[1]._fused_into() is c.main[6]
    assert m.main[6]._extra is m.res[1] and m.res[1]._fused_into() is m.main[6]
