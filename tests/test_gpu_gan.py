"""GPU parity: the tokenizer's auxiliary objectives — FrameDiscriminator, GANLoss (hinge) and
PerceptualLoss (VGG16 feature distance) against the real reference (oracle/make_golden.py:gen_gan_perceptual; frame picks
injected, VGG16 with closed-form weights because `weights='DEFAULT'` needs a download)."""
import pytest
import torch

from helpers import assert_close, det_weights, rel_l2
from oracle import fixtures as fx
from oracle import genie_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _grads(m):
    return {k: p.grad.float().cpu() for k, p in m.named_parameters() if p.grad is not None}


def _check_grads(got, gold, tol, what):
    assert set(got) == set(gold['norm']), (set(got) ^ set(gold['norm']))
    for k, n in gold['norm'].items():
        if n < 1e-8:
            continue
        assert abs(got[k].norm().item() - n) / n < tol, (what, k, got[k].norm().item(), n)
        if k in gold['full']:
            assert rel_l2(got[k], gold['full'][k]) < 2 * tol, (what, k, rel_l2(got[k], gold['full'][k]))


def test_frame_discriminator_against_reference(golden):
    from open_genie_b200.module.discriminator import FrameDiscriminator
    g = golden('gan_perceptual.pt')['disc']
    disc = FrameDiscriminator(inp_size=32)
    det_weights(disc)
    disc.to(DEV)
    frames = O.det_uniform('gan.frames', (4, 3, 32, 32)).to(DEV).requires_grad_(True)
    score = disc(frames)
    assert score.shape == (4,) and score.dtype == torch.float32
    (-score.mean()).backward()
    assert_close(score, g['score'], 3e-2, 3e-2 * g['score'].abs().max().item(), 'critic scores')
    # LeakyReLU's derivative jumps from 0.01 to 1 at zero: a pre-activation that bf16 noise moves across zero changes a
    # gradient path by two orders of magnitude, so gradients of this critic are noisier than those of the SiLU networks
    assert rel_l2(frames.grad.cpu(), g['dframes']) < 0.12
    _check_grads(_grads(disc), g['grads'], 0.1, 'critic')


def test_video_discriminator_against_reference(golden):
    """gan_discriminate='video': Conv3d stem, LeakyReLU residual blocks with blur-pool down-sampling, Linear head."""
    from open_genie_b200.module.discriminator import VideoDiscriminator
    g = golden('gan_perceptual.pt')['video_disc']
    vd = VideoDiscriminator(inp_size=(8, 32, 32))
    det_weights(vd)
    vd.to(DEV)
    clip = O.det_uniform('gan.clip', (2, 3, 8, 32, 32)).to(DEV).requires_grad_(True)
    score = vd(clip)
    assert score.shape == (2,) and score.dtype == torch.float32
    (-score.mean()).backward()
    assert_close(score, g['score'], 3e-2, 3e-2 * g['score'].abs().max().item(), 'video critic scores')
    assert rel_l2(clip.grad.cpu(), g['dclip']) < 0.12
    _check_grads(_grads(vd), g['grads'], 0.1, 'video critic')


def test_gan_hinge_losses_against_reference(golden):
    from open_genie_b200.module.loss import GANLoss
    g = golden('gan_perceptual.pt')['gan']
    gan = GANLoss(discriminate='frames', num_frames=2, inp_size=32)
    det_weights(gan.disc)
    gan.to(DEV)
    rec = O.det_uniform('gan.rec', (2, 3, 8, 32, 32)).to(DEV).requires_grad_(True)
    inp = O.det_uniform('gan.inp', (2, 3, 8, 32, 32)).to(DEV)
    idxs = g['frames_idxs'].to(DEV)
    gen_loss = gan(rec, inp, train_gen=True, frames_idxs=idxs)
    dis_loss = gan(rec, inp, train_gen=False, frames_idxs=idxs)
    (gen_loss + dis_loss).backward()
    assert abs(gen_loss.item() - g['gen_loss'].item()) < 3e-2 * abs(g['gen_loss'].item())
    assert abs(dis_loss.item() - g['dis_loss'].item()) < 3e-2 * abs(g['dis_loss'].item())
    assert rel_l2(rec.grad.cpu(), g['drec']) < 0.12                      # only the generator term reaches the video
    _check_grads(_grads(gan), g['grads'], 0.1, 'gan')
    # default path: random frame picks, same shapes
    assert gan(rec, inp, train_gen=True).dim() == 0


def test_perceptual_loss_against_reference(golden):
    from open_genie_b200.module.loss import PerceptualLoss
    g = golden('gan_perceptual.pt')['perc']
    perc = PerceptualLoss(num_frames=2)
    shapes = {k: tuple(v.shape) for k, v in perc.percept_model.state_dict().items()}
    perc.load_vgg_state_dict(O.det_state_dict(shapes, gain=1.4))
    perc.to(DEV)
    rec = O.det_uniform('gan.rec', (2, 3, 8, 32, 32)).to(DEV).requires_grad_(True)
    inp = O.det_uniform('gan.inp', (2, 3, 8, 32, 32)).to(DEV)
    loss = perc(rec, inp, frames_idxs=g['frames_idxs'].to(DEV))
    assert not loss.requires_grad                                            # the reference's probe detaches (misc.py:61)
    err = abs(loss.item() - g['loss'].item()) / g['loss'].item()
    assert err < 5e-2, (loss.item(), g['loss'].item())


def test_tokenizer_training_step_with_all_loss_terms():
    """VideoTokenizer with its constructor defaults (GAN + perceptual on): the step runs, the loss is the weighted sum of
    its logged terms, the critic trains, the perceptual extractor stays frozen."""
    import open_genie_b200 as og
    torch.manual_seed(0)
    tok = og.VideoTokenizer(fx.MINI_ENC, fx.MINI_DEC, d_codebook=fx.MINI_D_CODEBOOK, disc_kwargs={'inp_size': 32}).to(DEV)
    tok.train()
    opt = tok.configure_optimizers()
    video = torch.randn(2, 3, 8, 32, 32, device=DEV)
    with pytest.warns(UserWarning, match='no VGG16 weights'):
        loss, (rec, gen, dis, perc, ql) = tok(video)
    total = rec + gen + dis + perc + ql
    assert abs(loss.item() - total.item()) < 1e-4 * abs(total.item())
    loss.backward()
    named = dict(tok.named_parameters())
    assert named['gan_crit.disc.to_logits.3.weight'].grad is not None and named['enc_layers.0.conv3d.weight'].grad is not None
    assert all(p.grad is None for p in tok.perc_crit.parameters())
    before = named['gan_crit.disc.proj_in.weight'].detach().clone()
    opt.step()
    assert not torch.equal(before, named['gan_crit.disc.proj_in.weight'].detach())
    l2 = tok.training_step(video, 0)
    assert torch.isfinite(l2) and {'train_gen_loss', 'train_dis_loss', 'train_perc_loss'} <= set(tok.logged)
