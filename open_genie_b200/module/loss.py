"""Auxiliary objectives of the tokenizer — mirrors genie/module/loss.py:34-164 (PerceptualLoss, GANLoss)."""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib, ops
from .discriminator import FrameDiscriminator, VideoDiscriminator
from .image import Conv2dParams


def pick_frames(video: Tensor, frames_idxs: Tensor) -> Tensor:
    """genie/utils.py:30-56 with explicit indices: video (b, c, t, h, w) -> (b * k, c, h, w), frame frames_idxs[i] of clip
    i // k. Works on either tensor format (plain torch indexing; autograd flows back into `video`)."""
    b = video.shape[0]
    batch_idxs = torch.repeat_interleave(torch.arange(b, device=video.device), frames_idxs.numel() // b)
    return video[batch_idxs, :, frames_idxs.to(video.device)]


def random_frame_idxs(b: int, t: int, k: int, device) -> Tensor:
    """torch.cat([randperm(t)[:k] for _ in range(b)]) — loss.py:79-83 / 131-135."""
    return torch.cat([torch.randperm(t, device=device)[:k] for _ in range(b)])


# torchvision vgg16.features: ('C', out_channels) conv3x3+ReLU, 'M' max-pool; index = position in nn.Sequential
_VGG16_CFG = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M')


class _VGG16Features(nn.Module):
    """Parameter holder with torchvision's key names ('features.N.weight' / '.bias' for the conv at position N)."""

    def __init__(self) -> None:
        super().__init__()
        layers, cin = [], 3
        for v in _VGG16_CFG:
            if v == 'M':
                layers.append(nn.Identity())                   # MaxPool2d slot
            else:
                layers += [Conv2dParams(cin, v, 3), nn.Identity()]      # Conv2d + ReLU slot
                cin = v
        self.features = nn.Sequential(*layers)


class PerceptualLoss(nn.Module):
    """Mean feature-space MSE between VGG16 activations of randomly picked reconstructed / input frames —
    genie/module/loss.py:34-107. Two facts of the reference shape this implementation:
      * its RecordingProbe stores `out.clone().detach()` (genie/module/misc.py:61), so the loss carries NO gradient —
        it is a monitored value added to the objective; the whole pass therefore runs under no_grad;
      * `feat_layers` name ReLU outputs ('features.6', '.13', '.18', '.25' by default).
    Weights: torchvision's `weights='DEFAULT'` needs a download, which this offline image cannot do. Load a torchvision
    VGG16 state_dict with `load_vgg_state_dict(sd)` (keys 'features.N.{weight,bias}'); until then the extractor keeps its
    random initialisation and the value is only a smoke signal (a warning says so once)."""

    def __init__(self, model_name: str = 'vgg16', model_weights: str | None = 'DEFAULT', num_frames: int = 4,
                 feat_layers: str | Iterable[str] = ('features.6', 'features.13', 'features.18', 'features.25')) -> None:
        super().__init__()
        if model_name != 'vgg16':
            raise NotImplementedError('PerceptualLoss: only the reference default vgg16 extractor is implemented')
        self.num_frames = num_frames
        self.percept_model = _VGG16Features().requires_grad_(False)
        self.feat_layers = (feat_layers,) if isinstance(feat_layers, str) else tuple(feat_layers)
        idx = sorted(int(n.split('.')[1]) for n in self.feat_layers if n.startswith('features.'))
        assert len(idx) > 0, 'No valid layers found in the perceptual model.'
        self._taps, self._last = set(idx), max(idx)
        self.weights_loaded = False
        self._warned = False

    def load_vgg_state_dict(self, sd) -> None:
        own = self.percept_model.state_dict()
        self.percept_model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=True)
        self.weights_loaded = True

    @torch.no_grad()
    def features(self, frames: Tensor):
        """frames (n, 3, h, w) -> {layer index: internal bf16 activation} for the requested ReLU layers."""
        x = frames if frames.dim() == 5 else frames.unsqueeze(2)
        out = {}
        for i, layer in enumerate(self.percept_model.features):
            if i > self._last:
                break
            if isinstance(layer, Conv2dParams):
                x = ops.conv3d(x, layer.weight, layer.bias, layer.packed(), layer.geom)
            elif _VGG16_CFG_AT[i] == 'R':
                x = ops.activation(x, 'relu')
                if i in self._taps:
                    out[i] = x
            else:                                               # 2x2 max pool
                n, c, t, h, w = x.shape
                xi = ops.to_internal(x, torch.bfloat16)
                y = ops.empty_internal(n, c, 1, h // 2, w // 2, torch.bfloat16, x.device)
                _lib.call('og_maxpool2x2', xi.data_ptr(), y.data_ptr(), n, h, w, c, ops._stream())
                x = y
        return out

    @torch.no_grad()
    def forward(self, rec_video: Tensor, inp_video: Tensor, frames_idxs: Tensor | None = None) -> Tensor:
        b, c, t, h, w = inp_video.shape
        if not self.weights_loaded and not self._warned:
            import warnings
            warnings.warn('PerceptualLoss: no VGG16 weights loaded (offline image) — using the random initialisation')
            self._warned = True
        if frames_idxs is None:
            frames_idxs = random_frame_idxs(b, t, self.num_frames, inp_video.device)
        fake = self.features(pick_frames(rec_video.detach(), frames_idxs))
        real = self.features(pick_frames(inp_video, frames_idxs))
        terms = []
        for k in fake:
            a, r = fake[k], real[k]
            acc = torch.zeros((), dtype=torch.float32, device=a.device)
            _lib.call('og_sqdiff_sum', a.data_ptr(), r.data_ptr(), a.numel(), acc.data_ptr(), ops._stream())
            terms.append(acc / a.numel())
        return torch.stack(terms).mean()


def _vgg_roles():
    roles, = [[]]
    for v in _VGG16_CFG:
        if v == 'M':
            roles.append('M')
        else:
            roles += ['C', 'R']
    return tuple(roles)


_VGG16_CFG_AT = _vgg_roles()


class GANLoss(nn.Module):
    """Hinge GAN objective on randomly picked frames — genie/module/loss.py:109-164. `forward(rec, inp, train_gen)`:
    generator term -D(fake).mean(); critic term (relu(1 + D(fake.detach())) + relu(1 - D(real))).mean()."""

    def __init__(self, discriminate: str = 'frames', num_frames: int = 4, **kwargs) -> None:
        super().__init__()
        assert discriminate in ('frames', 'video'), 'Invalid discriminator type. Must be either "frames" or "video".'
        self.disc = FrameDiscriminator(**kwargs) if discriminate == 'frames' else VideoDiscriminator(**kwargs)
        self.num_frames = num_frames
        self.discriminate = discriminate

    def get_examples(self, rec_video: Tensor, inp_video: Tensor, frames_idxs: Tensor | None = None) -> Tuple[Tensor, Tensor]:
        b, c, t, h, w = inp_video.shape
        if self.discriminate == 'video':
            return rec_video, inp_video
        if frames_idxs is None:
            frames_idxs = random_frame_idxs(b, t, self.num_frames, inp_video.device)
        return pick_frames(rec_video, frames_idxs), pick_frames(inp_video, frames_idxs)

    def forward(self, rec_video: Tensor, inp_video: Tensor, train_gen: bool, frames_idxs: Tensor | None = None) -> Tensor:
        fake, real = self.get_examples(rec_video, inp_video, frames_idxs)
        fake_score = self.disc(fake) if train_gen else self.disc(fake.detach())
        if train_gen:
            return -fake_score.mean()
        real_score = self.disc(real)
        return (torch.relu(1 + fake_score) + torch.relu(1 - real_score)).mean()
