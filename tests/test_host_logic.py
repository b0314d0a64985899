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
                 'depth2spacetime_upsample', 'spacetime_downsample', 'group_norm', 'adaptive_group_norm', 'silu'):
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


def test_tokenizer_rejects_out_of_scope_losses():
    with pytest.raises(NotImplementedError, match='GAN and perceptual'):
        og.VideoTokenizer(fx.MINI_ENC, fx.MINI_DEC, d_codebook=6)


def test_conv_geometry():
    g = ops.ConvGeom(128, 128, (3, 3, 3), causal=True)
    assert (g.pt, g.ph, g.pw, g.direct) == (2, 1, 1, True)                    # video.py:155: (kt-1) + (1 - stride)
    g = ops.ConvGeom(128, 128, (3, 3, 3), causal=False)
    assert (g.pt, g.direct) == (1, True)
    g = ops.ConvGeom(128, 128, (3, 3, 3), stride=(2, 2, 2))
    assert (g.pt, g.direct, g.out_dims(16, 32, 32)) == (1, False, (8, 16, 16))
    g = ops.ConvGeom(3, 128, (3, 3, 3))
    assert not g.direct and g.kpad == 128 and g.k_main == 81
    g = ops.ConvGeom(128, 128, (3, 3, 3), stride=(1, 2, 2))
    assert g.out_dims(16, 64, 64) == (16, 32, 32)


def test_shard_batch_partitions_clips():
    for world in (1, 2, 4, 8):
        spans = [shard_batch(32, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 32
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert [shard_batch(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_zero_arena_lifetime_contract():
    """ops._ZeroArena: first step only measures, later steps hand out zeroed views of one buffer; a view is
    recycled (and re-zeroed) by the first allocation after mark_step()."""
    import torch
    from open_genie_b200 import ops
    a = ops._ZeroArena()
    assert a.zeros((2, 3), torch.float32, 'cpu').data_ptr() != 0          # disabled: plain torch.zeros
    a.enabled = True
    x = a.zeros((3, 4), torch.float32, 'cpu')                               # step 1: nothing to carve from yet
    assert a.buf.get(torch.device('cpu')) is None and x.sum() == 0
    a.mark_step()
    y = a.zeros((3, 4), torch.float32, 'cpu')
    z = a.zeros((5,), torch.float64, 'cpu')
    assert y.untyped_storage().data_ptr() == z.untyped_storage().data_ptr()  # one buffer
    y += 1
    z += 2
    big = a.zeros((1 << 20,), torch.float32, 'cpu')                          # does not fit: falls back, grows next step
    assert big.untyped_storage().data_ptr() != y.untyped_storage().data_ptr()
    a.mark_step()
    w = a.zeros((3, 4), torch.float32, 'cpu')
    assert w.sum() == 0 and a.buf[torch.device('cpu')].numel() >= 4 << 20
