"""GPU parity: the inference path — DynamicsModel.generate (MaskGIT sampling, genie/dynamics.py:101-165) against the
REAL reference loop run with injected uniform draws (oracle/make_golden.py:gen_generate), Genie.forward roll-out
(genie/genie.py:65-105, re-specified), and the 'blur_pool' registry module (genie/module/image.py:43-85)."""
import pytest
import torch

from helpers import assert_close, det_weights
from oracle import fixtures as fx
from oracle import genie_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def test_maskgit_sampler_matches_reference_loop_bit_exact(golden):
    """Same logits, same uniforms, same schedule -> the fused sampler must reproduce the reference loop's tokens exactly
    (integer output). P = 64 / V = 64 / 5 steps from the mini model, and P = 256 / V = 1024 / 25 steps (configs[3] sizes)."""
    import open_genie_b200 as og
    from open_genie_b200 import ops
    g = golden('generate.pt')
    dm = og.DynamicsModel(fx.MINI_DYN_DESC, **fx.MINI_DYN)
    m = g['mini']
    code, mask = ops.maskgit_sample(m['logits_last'].to(DEV), m['uniforms'].to(DEV), dm.get_schedule(m['steps'], (8, 8)))
    assert int(mask.sum()) == 0
    assert torch.equal(code.cpu(), m['pred_tok'][:, -1])
    logits = O.det_uniform('gen.big.logits', (2, 16, 16, 1024), 2.0).to(DEV)
    uni = ((O.det_uniform('gen.big.uniforms', (25, 2 * 256)) / (3 ** 0.5) + 1) * 0.5).to(DEV)
    for which in ('linear', 'cosine'):
        ref = g[f'big_{which}']['pred_tok'][:, -1]
        code, mask = ops.maskgit_sample(logits, uni, dm.get_schedule(25, (16, 16), which))
        assert int(mask.sum()) == 0
        same = (code.cpu() == ref).float().mean().item()
        assert same == 1.0, (which, same)
    # the oracle restatement agrees with the same golden (ties the three implementations together)
    ref = O.maskgit_generate(logits.cpu(), g['big_linear']['tokens'], uni.cpu(), dm.get_schedule(25, (16, 16), 'linear'))
    assert torch.equal(ref, g['big_linear']['pred_tok'])


def test_generate_runs_the_transformer_once_and_keeps_history(golden):
    import open_genie_b200 as og
    from open_genie_b200 import _lib, ops
    g = golden('generate.pt')['mini']
    dm = og.DynamicsModel(fx.MINI_DYN_DESC, **fx.MINI_DYN)
    det_weights(dm)
    dm.to(DEV)
    tokens, act, uni = g['tokens'].to(DEV), g['act'].to(DEV), g['uniforms'].to(DEV)
    n0 = _lib.launch_count()
    out = dm.generate(tokens, act, steps=g['steps'], uniforms=uni)
    launches = _lib.launch_count() - n0
    assert out.shape == (2, 4, 8, 8) and out.dtype == tokens.dtype and torch.equal(out[:, :3], tokens)
    assert int(out.min()) >= 0 and int(out.max()) < fx.MINI_DYN['tok_vocab']
    # one transformer evaluation (2 blocks) + CDF + ONE sampling launch for all 5 iterations: far fewer launches than
    # five evaluations of the model would need
    assert launches < 40, launches
    # self-consistency: the same tokens as the sampler applied to this model's own last-frame logits
    tok_id = torch.cat([tokens, torch.zeros(2, 1, 8, 8, dtype=tokens.dtype, device=DEV)], 1)
    act_id = torch.cat([act, torch.zeros(2, 1, dtype=act.dtype, device=DEV)], 1)
    _, last = dm(tok_id, act_id)
    code, _ = ops.maskgit_sample(dm._logits(tok_id, act_id)[:, -1], uni, dm.get_schedule(g['steps'], (8, 8)))
    assert torch.equal(out[:, -1], code)
    # and close to the reference's own roll-out: the bf16 logits move a CDF boundary past a draw only rarely
    agree = (out.cpu() == g['pred_tok']).float().mean().item()
    assert agree > 0.9, agree
    # default path: draws from torch's CUDA generator
    out2 = dm.generate(tokens, act, steps=4)
    assert out2.shape == (2, 4, 8, 8) and torch.equal(out2[:, :3], tokens)


def test_genie_forward_rolls_out_a_video():
    """Genie.forward (inference): prompt -> tokens -> num_frames x generate -> decode."""
    import open_genie_b200 as og
    torch.manual_seed(0)
    no_time = lambda bp: tuple((n, {**kw, **({'time_factor': 1} if 'time_factor' in kw else {})}) for n, kw in bp)
    tok = og.VideoTokenizer(no_time(fx.MINI_ENC), no_time(fx.MINI_DEC), d_codebook=fx.MINI_D_CODEBOOK, gan_loss_weight=0,
                            perc_loss_weight=0)
    gen = og.Genie(tok,
                   dict(enc_desc=fx.MINI_ACT_ENC, dec_desc=fx.MINI_ACT_DEC, d_codebook=4, n_embd=128, inp_shape=(32, 32)),
                   dict(desc=fx.MINI_DYN_DESC, tok_vocab=2 ** fx.MINI_D_CODEBOOK, act_vocab=16, embed_dim=128)).to(DEV)
    prompt = torch.randn(2, 3, 32, 32, device=DEV)                        # image prompt (b, c, h, w)
    actions = torch.randint(0, 16, (2, 4), device=DEV)
    video = gen(prompt, actions, num_frames=3, steps_per_frame=4)
    assert video.shape == (2, 3, 4, 32, 32) and video.dtype == torch.float32 and torch.isfinite(video).all()
    # the prompt frame's tokens are kept: decoding only them gives the same first frame up to the causal decoder
    _, tok0 = gen.tokenizer.tokenize(prompt[:, :, None])
    assert tok0.reshape(2, -1, 8, 8).shape[1] == 1
    # codes_from_indices inverts the LFQ bit packing
    q, idx = gen.tokenizer.tokenize(torch.randn(2, 3, 2, 32, 32, device=DEV))
    assert torch.equal(gen.tokenizer.quant.codes_from_indices(idx).sign(), q.sign())
    with pytest.raises(ValueError, match='one action per generated transition'):
        gen(prompt, actions[:, :1], num_frames=3, steps_per_frame=2)


def test_blur_pool_registry_module_matches_reference(golden):
    """get_module('blur_pool') -> BlurPooling2d (genie/module/__init__.py:33), values vs the reference."""
    from open_genie_b200.module import get_module
    g = golden('blur2d.pt')
    cls = get_module('blur_pool')
    x = O.det_uniform('kat.blur2d.x', (2, 16, 12, 12)).to(torch.bfloat16).float().to(DEV).requires_grad_(True)
    for k, s in ((3, 2), (4, 2)):
        m = cls(k, stride=s).to(DEV)
        ref = g[f'k{k}s{s}']
        assert torch.equal(m.blur.cpu(), ref['blur'])
        y = m(x)
        y.backward(O.det_uniform(f'kat.blur2d.g.{k}', tuple(y.shape)).to(DEV))
        assert_close(y, ref['y'], 1e-2, 1e-2 * ref['y'].abs().max().item(), f'blur2d k{k} y')
        assert_close(x.grad, ref['dx'], 2e-2, 2e-2 * ref['dx'].abs().max().item(), f'blur2d k{k} dx')
        x.grad = None
