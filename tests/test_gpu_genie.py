"""GPU parity: LatentAction, DynamicsModel (forward / compute_loss / generate) and the composed Genie step,
against the golden vectors produced by the real reference in its HEAD-valid configurations."""
import pytest
import torch

from helpers import assert_close, det_weights, rel_l2
from oracle import fixtures as fx
from oracle import genie_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _grads(m):
    return {k: p.grad.float().cpu() for k, p in m.named_parameters() if p.grad is not None}


def test_latent_action_against_reference_golden(golden):
    import open_genie_b200 as og
    g = golden('latent_action_mini.pt')
    la = og.LatentAction(fx.MINI_ACT_ENC, fx.MINI_ACT_DEC, d_codebook=fx.MINI_ACT_D_CODEBOOK, n_embd=fx.MINI_ACT_EMBD,
                         inp_shape=fx.MINI_ACT_VIDEO_SHAPE[-2:])
    det_weights(la)
    la.to(DEV).train()
    video = O.det_uniform('action.video', fx.MINI_ACT_VIDEO_SHAPE).to(DEV)
    idxs, loss, (rec_loss, q_loss) = la(video)
    loss.backward()
    assert idxs.shape == g['idxs'].shape and idxs.dtype == torch.int64
    assert (idxs.cpu() == g['idxs']).float().mean() >= 0.75          # 4-bit codes of 8 frames; near-zero logits may flip
    assert abs(rec_loss.item() - g['rec_loss'].item()) / g['rec_loss'].item() < 3e-2
    assert abs(loss.item() - g['loss'].item()) / abs(g['loss'].item()) < 0.15     # q_loss carries beta=100 sensitivity
    grads = _grads(la)
    assert set(grads) == set(g['grads']['norm'])
    for k, n in g['grads']['norm'].items():
        if k.startswith(('dec_layers', 'proj_out')) and n > 1e-6:
            assert abs(grads[k].norm().item() - n) / n < 0.1, (k, grads[k].norm().item(), n)
        assert torch.isfinite(grads[k]).all(), k


def test_dynamics_against_reference_golden(golden):
    import open_genie_b200 as og
    g = golden('dynamics_mini.pt')
    dm = og.DynamicsModel(fx.MINI_DYN_DESC, **fx.MINI_DYN)
    det_weights(dm)
    dm.to(DEV)
    tokens, act, mask = g['tokens'].to(DEV), g['act'].to(DEV), g['mask'].to(DEV)
    logits, last = dm(tokens, act)
    assert logits.shape == g['logits'].shape and last.shape == g['logits'][:, -1].shape and logits.dtype == torch.float32
    assert rel_l2(logits.cpu(), g['logits']) < 2e-2
    loss = dm.compute_loss(tokens, act, mask=mask)
    loss.backward()
    assert abs(loss.item() - g['loss'].item()) / g['loss'].item() < 2e-2
    grads = _grads(dm)
    assert set(grads) == set(g['grads']['norm'])
    for k, v in g['grads']['full'].items():
        assert rel_l2(grads[k], v) < 0.1, (k, rel_l2(grads[k], v))
    for k, n in g['grads']['norm'].items():
        assert abs(grads[k].norm().item() - n) / n < 0.1, (k, grads[k].norm().item(), n)


def test_dynamics_generate_maskgit_shapes():
    import open_genie_b200 as og
    torch.manual_seed(0)
    dm = og.DynamicsModel(fx.MINI_DYN_DESC, **fx.MINI_DYN).to(DEV)
    tokens = torch.randint(0, fx.MINI_DYN['tok_vocab'], (2, 3, 8, 8), device=DEV)
    act = torch.randint(0, fx.MINI_DYN['act_vocab'], (2, 3), device=DEV)
    out = dm.generate(tokens, act, steps=5)
    assert out.shape == (2, 4, 8, 8) and torch.equal(out[:, :3], tokens)
    assert int(out.min()) >= 0 and int(out.max()) < fx.MINI_DYN['tok_vocab']
    assert dm.get_schedule(10, (16, 16), 'cosine').tolist() == [1, 7, 14, 21, 26, 32, 36, 39, 41, 39]


def test_genie_training_step_composition():
    import open_genie_b200 as og
    torch.manual_seed(0)
    # Genie needs token frames == action frames (tok_emb + act_emb broadcast, genie/dynamics.py:55): use the mini
    # tokenizer blueprint without its temporal compression
    no_time = lambda bp: tuple((n, {**kw, **({'time_factor': 1} if 'time_factor' in kw else {})}) for n, kw in bp)
    tok = og.VideoTokenizer(no_time(fx.MINI_ENC), no_time(fx.MINI_DEC), d_codebook=fx.MINI_D_CODEBOOK, gan_loss_weight=0,
                            perc_loss_weight=0)
    gen = og.Genie(tok,
                   dict(enc_desc=fx.MINI_ACT_ENC, dec_desc=fx.MINI_ACT_DEC, d_codebook=4, n_embd=128, inp_shape=(32, 32)),
                   dict(desc=fx.MINI_DYN_DESC, tok_vocab=2 ** fx.MINI_D_CODEBOOK, act_vocab=16, embed_dim=128)).to(DEV)
    opt = gen.configure_optimizers()
    video = torch.randn(2, 3, 8, 32, 32, device=DEV)
    losses = []
    for _ in range(3):
        loss = gen.training_step(video, 0)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        losses.append(loss.item())
    assert all(l == l for l in losses)
    assert all(p.grad is None for p in gen.tokenizer.parameters())            # frozen tokenizer
    assert {'train_loss', 'train/act_loss', 'train/dyn_loss', 'train/act_rec_loss', 'train/act_q_loss'} <= set(gen.logged)
    assert abs(gen.logged['train_loss'].item() - (gen.logged['train/act_loss'] + gen.logged['train/dyn_loss']).item()) < 1e-3
