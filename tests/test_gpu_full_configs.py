"""GPU parity on BASELINE.json's FULL-SIZE configurations against golden vectors produced by the real reference
(oracle/make_golden_full.py, B = 2 clips of 3x16x64x64):

  configs[0]  tokenize() + decode(), MAGVIT2_ENC/DEC                       test_cfg0_*
  configs[1]  VideoTokenizer training step (loss terms + every gradient)   test_cfg1_*
  configs[2]  LatentAction, pinned LATENT_ACT blueprints at 64x64          test_cfg2_*   (S = 4096 flash attention,
              T = 16 temporal attention, 2-D RoPE at 4096 positions, 262144 -> 8 to_act projection)
  configs[3]  DynamicsModel d_model = 512, 8 heads, 16x16x16 tokens        test_cfg3_*
  configs[4]  Genie training step (frozen tokenizer + action + dynamics)   test_cfg4_*

Tolerances. Integer outputs (token / action indices) must equal the reference wherever the pre-sign value is
further from zero than the bf16 error of the chain that produced it (stated per test). "chain" tests run the same
rounding points as the product (bf16-rounded operands, no sign quantiser in the differentiated path) and hold EVERY
parameter gradient to a relative L2 error on a fixed sample; "fp32" tests compare with the reference exactly as a
user runs it and carry the looser, documented end-to-end tolerances of DESIGN.md §1."""
import os

import pytest
import torch

from helpers import det_weights, rel_l2
from oracle import fixtures as fx
from oracle import genie_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'
VERBOSE = os.environ.get('OG_TEST_VERBOSE', '0') != '0'


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def sample(p_grad, key):
    g = p_grad.detach().float().flatten()
    return g[O.det_indices(key, g.numel()).to(g.device)].cpu()


def check_grads(module, gold, prefixes, tol_l2, tol_norm, what, floor=1e-7):
    """Every parameter whose name starts with one of `prefixes`: gradient norm within tol_norm and relative L2 error
    on the golden's deterministic sample within tol_l2. Returns the worst offenders for the log."""
    worst_l2, worst_n = (0.0, None), (0.0, None)
    named = dict(module.named_parameters())
    checked = 0
    for k, n in gold['norm'].items():
        if not k.startswith(prefixes):
            continue
        p = named[k]
        assert p.grad is not None, f'{what}: no gradient for {k}'
        assert torch.isfinite(p.grad).all(), f'{what}: non-finite gradient for {k}'
        if n < floor:
            continue
        got_n = p.grad.float().norm().item()
        en = abs(got_n - n) / n
        el = rel_l2(sample(p.grad, k), gold['sample'][k])
        if el > worst_l2[0]:
            worst_l2 = (el, k)
        if en > worst_n[0]:
            worst_n = (en, k)
        checked += 1
    if VERBOSE:
        print(f'[{what}] {checked} gradients; worst rel-L2 {worst_l2[0]:.3e} ({worst_l2[1]}); '
              f'worst norm error {worst_n[0]:.3e} ({worst_n[1]})')
    assert checked > 0
    assert worst_l2[0] < tol_l2, (what, worst_l2)
    assert worst_n[0] < tol_norm, (what, worst_n)


def bits_of(idxs, d):
    return ((idxs.cpu()[..., None] >> torch.arange(d - 1, -1, -1)) & 1).bool()


# ------------------------------------------------------------------------------------------------
# configs[0] / configs[1]: MAGVIT2 VideoTokenizer
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def full_tok():
    import open_genie_b200 as og
    tok = og.VideoTokenizer(og.MAGVIT2_ENC_DESC, og.MAGVIT2_DEC_DESC, d_codebook=fx.FULL_D_CODEBOOK, gan_loss_weight=0,
                            perc_loss_weight=0)
    sd = det_weights(tok)
    return tok.to(DEV), sd


def test_cfg0_tokenize_decode_full_magvit2(golden, full_tok):
    g = golden('full_tokenizer.pt')
    tok, _ = full_tok
    assert sum(p.numel() for p in tok.parameters()) == g['n_params'] == 375554837
    video = O.det_uniform('full.tokenizer.video', fx.FULL_VIDEO_SHAPE).to(DEV)
    quant, idxs = tok.tokenize(video)
    assert quant.shape == (2, 18, 4, 8, 8) and idxs.shape == (2, 4, 8, 8) and idxs.dtype == torch.int64
    # 33 bf16 layers precede the sign. The latent itself agrees with the reference to ~1.1e-2 relative L2
    # (test_cfg1_encoder_chain), i.e. a per-element error of ~1.5 % of the mean magnitude: a bit may differ only where
    # |latent| lies within a few of those errors of zero (the 10 % band); everywhere else the packed bits must be equal
    enc = g['enc'].movedim(1, -1)
    ref_bits, got_bits = bits_of(g['idxs'], 18), bits_of(idxs, 18)
    band = 0.10 * enc.abs().mean()
    safe = enc.abs() > band
    agree = (ref_bits == got_bits).float().mean().item()
    diff = ref_bits != got_bits
    worst = (enc.abs()[diff].max() / enc.abs().mean()).item() if diff.any() else 0.0
    if VERBOSE:
        print(f'[cfg0] bit agreement {agree:.4f} ({int(diff.sum())} of {diff.numel()} differ; largest |latent| among them = '
              f'{worst:.4f} x mean); safe fraction {safe.float().mean().item():.4f}')
    assert torch.equal(ref_bits[safe], got_bits[safe]), f'index bits differ on well-separated latents ({worst:.3f} x mean)'
    assert agree > 0.98, agree
    assert torch.equal(quant.cpu().sign()[:, :, :][safe.movedim(-1, 1)], g['quant'].float()[safe.movedim(-1, 1)])
    dec = tok.decode(g['quant'].float().to(DEV))           # decode the REFERENCE's codes
    assert dec.shape == fx.FULL_VIDEO_SHAPE and dec.dtype == torch.float32 and dec.is_contiguous()
    err = rel_l2(dec.cpu(), g['decode'].float())
    if VERBOSE:
        print(f'[cfg0] decode rel-L2 {err:.3e}')
    assert err < 3e-2, err


def test_cfg1_encoder_chain_full_magvit2(golden, full_tok):
    """Encoder forward + backward for a FIXED upstream gradient: all 3x3x3 / 1x1x1 / strided convs, the 512-channel
    4x8x8 split-K stage, fused GroupNorm statistics — every encoder gradient against the reference chain run."""
    from open_genie_b200 import ops
    g = golden('full_tokenizer.pt')['chain']
    tok, _ = full_tok
    tok.train()
    tok.zero_grad(set_to_none=True)
    video = bf16_round(O.det_uniform('full.tokenizer.video', fx.FULL_VIDEO_SHAPE)).to(DEV)
    enc = tok.encode(video)
    gup = O.det_uniform('full.enc.upstream', tuple(enc.shape)).to(DEV)
    enc.backward(gup.to(enc.dtype).permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3))
    err = rel_l2(ops.to_reference(enc).cpu(), g['enc'])
    if VERBOSE:
        print(f'[cfg1 enc chain] latent rel-L2 {err:.3e}')
    assert err < 2e-2, err
    check_grads(tok, g['enc_grads'], ('enc_layers',), 8e-2, 5e-2, 'cfg1 encoder chain')


def test_cfg1_decoder_chain_full_magvit2(golden, full_tok):
    """Decoder forward + backward from the reference's codes: the 512->4096 / 256->2048 / 256->1024 up-convolutions +
    pixel shuffles, AdaGN, the operand-swapped C=128 GEMMs, the 128->3 tail — every decoder gradient."""
    from open_genie_b200 import ops
    gt = golden('full_tokenizer.pt')
    g = gt['chain']
    tok, _ = full_tok
    tok.train()
    tok.zero_grad(set_to_none=True)
    video = bf16_round(O.det_uniform('full.tokenizer.video', fx.FULL_VIDEO_SHAPE)).to(DEV)
    rec = tok._decode_internal(ops.to_internal(gt['quant'].float().to(DEV), torch.float32))
    loss = ops.mse_loss(rec, video)
    loss.backward()
    err = rel_l2(ops.to_reference(rec).cpu(), g['rec'].float())
    el = abs(loss.item() - g['rec_loss'].item()) / g['rec_loss'].item()
    if VERBOSE:
        print(f'[cfg1 dec chain] reconstruction rel-L2 {err:.3e}, rec loss rel err {el:.3e}')
    assert err < 3e-2 and el < 1e-2, (err, el)
    check_grads(tok, g['dec_grads'], ('dec_layers',), 8e-2, 5e-2, 'cfg1 decoder chain')


def test_cfg1_training_step_full_magvit2(golden, full_tok):
    """The user-visible training forward/backward against the fp32 reference run (end-to-end tolerances)."""
    g = golden('full_tokenizer.pt')
    tok, _ = full_tok
    tok.train()
    tok.zero_grad(set_to_none=True)
    video = O.det_uniform('full.tokenizer.video', fx.FULL_VIDEO_SHAPE).to(DEV)
    loss, (rec, gen, dis, perc, ql) = tok(video)
    loss.backward()
    assert gen == 0 and dis == 0 and perc == 0
    er = abs(rec.item() - g['rec_loss'].item()) / g['rec_loss'].item()
    eq = abs(ql.item() - g['quant_loss'].item()) / abs(g['quant_loss'].item())
    et = abs(loss.item() - g['loss'].item()) / g['loss'].item()
    if VERBOSE:
        print(f'[cfg1 step] rec {rec.item():.6f} vs {g["rec_loss"].item():.6f} ({er:.2e}); quant {ql.item():.6f} vs '
              f'{g["quant_loss"].item():.6f} ({eq:.2e}); loss ({et:.2e})')
    assert er < 2e-2 and eq < 5e-2 and et < 3e-2, (er, eq, et)
    grads = {k for k, p in tok.named_parameters() if p.grad is not None}
    assert grads == set(g['grads']['norm'])
    # decoder gradients: a few of the 9216 code bits differ from the fp32 run (see cfg0), which flips +-1 inputs of the
    # decoder — the first 512-channel blocks at 4x8x8 feel that most (measured 0.23 relative L2; the same decoder on
    # IDENTICAL codes is held to 8e-2 per gradient in test_cfg1_decoder_chain). Scale and direction must still agree.
    check_grads(tok, g['grads'], ('dec_layers',), 0.35, 0.1, 'cfg1 step, decoder vs fp32 reference')
    # encoder gradients pass through d/dx of the LFQ entropy at beta = 100 (width 0.0025 in x): scale-gated only
    named = dict(tok.named_parameters())
    for k, n in g['grads']['norm'].items():
        if k.startswith('enc_layers') and n > 1e-6:
            r = named[k].grad.float().norm().item() / n
            assert 0.5 < r < 2.0, (k, r)


# ------------------------------------------------------------------------------------------------
# configs[2]: LatentAction on the pinned LATENT_ACT blueprints at 64x64
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def full_action():
    import open_genie_b200 as og
    la = og.LatentAction(og.LATENT_ACT_ENC, og.LATENT_ACT_DEC, d_codebook=fx.FULL_ACT_D_CODEBOOK, n_embd=fx.FULL_ACT_EMBD,
                         inp_shape=fx.FULL_VIDEO_SHAPE[-2:])
    assert og.LATENT_ACT_ENC == fx.FULL_ACT_ENC and og.LATENT_ACT_DEC == fx.FULL_ACT_DEC
    det_weights(la)
    return la.to(DEV).train()


def _action_logits(la, video):
    x = la.proj_in(video)
    for enc in la.enc_layers:
        x = enc(x)
    return la.to_act(x), x


def test_cfg2_latent_action_chain(golden, full_action):
    g = golden('full_latent_action.pt')['chain']
    la = full_action
    la.zero_grad(set_to_none=True)
    from open_genie_b200 import ops
    video = bf16_round(O.det_uniform('full.action.video', fx.FULL_VIDEO_SHAPE)).to(DEV)
    logits, enc_video = _action_logits(la, video)
    assert logits.shape == (2, 16, 8) and logits.dtype == torch.float32
    q_fixed = O.det_uniform('full.action.qfixed', (2, 16, 8)).sign().to(DEV)
    g_fixed = O.det_uniform('full.action.glogits', (2, 16, 8), 1e-3).to(DEV)
    recon = la.decode(enc_video, q_fixed)
    rl = ops.mse_loss(recon, video)
    (rl + (logits * g_fixed).sum()).backward()
    e_log = rel_l2(logits.cpu(), g['logits'])
    e_rec = rel_l2(ops.to_reference(recon).cpu(), g['recon'].float())
    e_rl = abs(rl.item() - g['rec_loss'].item()) / g['rec_loss'].item()
    if VERBOSE:
        print(f'[cfg2 chain] logits rel-L2 {e_log:.3e}; recon rel-L2 {e_rec:.3e}; rec loss {e_rl:.3e}')
    assert e_log < 3e-2 and e_rec < 3e-2 and e_rl < 2e-2, (e_log, e_rec, e_rl)
    check_grads(la, g['grads'], ('proj_in', 'proj_out', 'enc_layers', 'dec_layers', 'to_act'), 0.1, 6e-2, 'cfg2 chain')


def test_cfg2_latent_action_step(golden, full_action):
    g = golden('full_latent_action.pt')
    la = full_action
    la.zero_grad(set_to_none=True)
    assert sum(p.numel() for p in la.parameters()) == g['n_params'] == 25174019
    video = O.det_uniform('full.action.video', fx.FULL_VIDEO_SHAPE).to(DEV)
    idxs, loss, (rec_loss, q_loss) = la(video)
    loss.backward()
    assert idxs.shape == (2, 16) and idxs.dtype == torch.int64
    # action bits: equal wherever the reference's logit is outside the chain's bf16 error band
    logits = g['logits']
    ref_bits, got_bits = bits_of(g['idxs'], 8), bits_of(idxs, 8)
    safe = logits.abs() > 0.05 * logits.abs().mean()
    assert torch.equal(ref_bits[safe], got_bits[safe]), 'action bits differ on well-separated logits'
    agree = (ref_bits == got_bits).float().mean().item()
    er = abs(rec_loss.item() - g['rec_loss'].item()) / g['rec_loss'].item()
    el = abs(loss.item() - g['loss'].item()) / abs(g['loss'].item())
    if VERBOSE:
        print(f'[cfg2 step] bit agreement {agree:.4f} (safe {safe.float().mean().item():.3f}); rec loss {er:.2e}; loss {el:.2e}')
    assert agree >= 0.95 and er < 3e-2, (agree, er)
    if agree == 1.0:
        assert el < 5e-2, el
    assert {k for k, p in la.named_parameters() if p.grad is not None} == set(g['grads']['norm'])
    if agree == 1.0:     # identical codes: the decoder sees the same conditioning as the fp32 run
        check_grads(la, g['grads'], ('dec_layers', 'proj_out'), 0.15, 0.1, 'cfg2 step, decoder vs fp32 reference')


# ------------------------------------------------------------------------------------------------
# configs[3]: DynamicsModel, d_model = 512, 8 heads
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def full_dyn():
    import open_genie_b200 as og
    dm = og.DynamicsModel(fx.FULL_DYN_DESC, **fx.FULL_DYN)
    det_weights(dm)
    return dm.to(DEV)


def test_cfg3_dynamics_full(golden, full_dyn):
    g = golden('full_dynamics.pt')
    dm = full_dyn
    dm.zero_grad(set_to_none=True)
    tokens, act, mask = g['tokens'].to(DEV), g['act'].to(DEV), g['mask'].to(DEV)
    logits, last = dm(tokens, act)
    assert logits.shape == (2, 16, 16, 16, 1024) and logits.dtype == torch.float32 and last.shape == (2, 16, 16, 1024)
    sub = logits[:, ::4, ::4, ::4].cpu()
    e32, ech = rel_l2(sub, g['logits_sub']), rel_l2(sub, g['chain']['logits_sub'])
    en = abs(logits.norm().item() - g['logits_norm'].item()) / g['logits_norm'].item()
    loss = dm.compute_loss(tokens, act, mask=mask)
    loss.backward()
    el32 = abs(loss.item() - g['loss'].item()) / g['loss'].item()
    elch = abs(loss.item() - g['chain']['loss'].item()) / g['chain']['loss'].item()
    if VERBOSE:
        print(f'[cfg3] logits rel-L2 vs fp32 {e32:.3e}, vs rounded-operand run {ech:.3e}; norm {en:.2e}; '
              f'loss {el32:.2e} / {elch:.2e}')
    assert e32 < 2e-2 and ech < 2e-2 and en < 1e-2 and el32 < 1e-2 and elch < 1e-2
    assert {k for k, p in dm.named_parameters() if p.grad is not None} == set(g['grads']['norm'])
    check_grads(dm, g['chain']['grads'], ('',), 8e-2, 5e-2, 'cfg3 vs rounded-operand reference run')
    check_grads(dm, g['grads'], ('',), 0.1, 6e-2, 'cfg3 vs fp32 reference run')


# ------------------------------------------------------------------------------------------------
# configs[4]: Genie — frozen tokenizer (REPR_TOK, no temporal compression) + LatentAction + DynamicsModel
# ------------------------------------------------------------------------------------------------
def _repr_bp(desc):
    return tuple((n, {**kw, **({'n_rep': fx.GENIE_TOK_N_REP} if 'n_rep' in kw else {})}) for n, kw in desc)


def test_cfg4_genie_training_step_vs_reference_composition(golden, full_action, full_dyn):
    import open_genie_b200 as og
    g = golden('full_genie.pt')
    ga = golden('full_latent_action.pt')
    tok = og.VideoTokenizer(_repr_bp(og.REPR_TOK_ENC), _repr_bp(og.REPR_TOK_DEC), d_codebook=fx.GENIE_TOK_D_CODEBOOK,
                            gan_loss_weight=0, perc_loss_weight=0)
    det_weights(tok)
    genie = og.Genie(tok, full_action, full_dyn).to(DEV)
    genie.zero_grad(set_to_none=True)
    video = O.det_uniform('full.action.video', fx.FULL_VIDEO_SHAPE).to(DEV)
    # frozen tokenizer: token grid (b, 16, 16, 16), vocabulary 1024 — bits equal outside the bf16 error band
    quant, tokens = genie.tokenizer.tokenize(video)
    assert tokens.shape == (2, 16, 16, 16) and quant.shape == (2, 512, 16, 16, 16)
    pre = g['pre_sign']
    ref_bits, got_bits = bits_of(g['tokens'], 10), bits_of(tokens, 10)
    safe = pre.abs() > 0.10 * pre.abs().mean()
    agree = (ref_bits == got_bits).float().mean().item()
    diff = ref_bits != got_bits
    worst = (pre.abs()[diff].max() / pre.abs().mean()).item() if diff.any() else 0.0
    if VERBOSE:
        print(f'[cfg4] token bits: {int(diff.sum())} of {diff.numel()} differ; largest |pre-sign| among them {worst:.4f} x mean')
    assert torch.equal(ref_bits[safe], got_bits[safe]), f'token bits differ on well-separated latents ({worst:.3f} x mean)'
    dec = genie.tokenizer.decode(quant)
    assert dec.shape == fx.FULL_VIDEO_SHAPE
    # the composed step (genie/genie.py:107-125) with the golden's mask
    mask = g['mask'].to(DEV)
    loss, aux = genie.compute_loss(video, mask=mask)
    loss.backward()
    aux = dict(aux)
    assert all(p.grad is None for p in genie.tokenizer.parameters())
    assert abs(loss.item() - (aux['act_loss'] + aux['dyn_loss']).item()) < 1e-4 * abs(loss.item())
    e_rec = abs(aux['act_rec_loss'].item() - ga['rec_loss'].item()) / ga['rec_loss'].item()
    # dynamics term, isolated: the reference's own tokens and action ids through our DynamicsModel
    genie.dynamics_model.zero_grad(set_to_none=True)
    dyn = genie.dynamics_model.compute_loss(g['tokens'].to(DEV), g['act_id'].to(DEV), mask=mask)
    dyn.backward()
    e_dyn = abs(dyn.item() - g['dyn_loss'].item()) / g['dyn_loss'].item()
    e_dyn_step = abs(aux['dyn_loss'].item() - g['dyn_loss'].item()) / g['dyn_loss'].item()
    if VERBOSE:
        print(f'[cfg4] token bit agreement {agree:.4f}; act rec loss {e_rec:.2e}; dyn loss on reference tokens {e_dyn:.2e}; '
              f'dyn loss in the composed step {e_dyn_step:.2e}')
    assert agree > 0.97 and e_rec < 3e-2 and e_dyn < 1e-2 and e_dyn_step < 5e-2
    check_grads(genie.dynamics_model, g['dyn_grads'], ('',), 0.1, 6e-2, 'cfg4 dynamics gradients on reference tokens')


def test_cfg4_genie_rejects_misaligned_token_and_action_frames():
    """With a time-compressing tokenizer the reference's tok_emb + act_emb broadcast raises (dynamics.py:55); so do we."""
    import open_genie_b200 as og
    dm = og.DynamicsModel(fx.MINI_DYN_DESC, **fx.MINI_DYN).to(DEV)
    tokens = torch.zeros(2, 4, 8, 8, dtype=torch.int64, device=DEV)
    with pytest.raises(ValueError, match='does not match'):
        dm(tokens, torch.zeros(2, 8, dtype=torch.int64, device=DEV))
