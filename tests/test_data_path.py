"""Data path: Platformer2D / LightningPlatformer2D against the reference's own classes on mp4 files
written here with OpenCV (CPU), and the device-side frame decode + prefetcher (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def clips(tmp_path_factory):
    cv2 = pytest.importorskip('cv2')
    root = tmp_path_factory.mktemp('platformer')
    rng = np.random.default_rng(0)
    for split, n in (('train', 5), ('val', 2), ('test', 2)):
        d = root / 'Coinrun' / split
        d.mkdir(parents=True)
        for i in range(n):
            w = cv2.VideoWriter(str(d / f'clip{i}.mp4'), cv2.VideoWriter_fourcc(*'mp4v'), 15, (64, 64))
            assert w.isOpened()
            base = rng.integers(0, 255, (8, 8, 3))
            for t in range(12 if i == 0 else 20):          # clip0 is SHORTER than num_frames = 16
                frame = np.kron((base + 9 * t) % 256, np.ones((8, 8, 1))).astype(np.uint8)
                w.write(frame)
            w.release()
    return str(root)


def _reference_classes():
    """The reference's own dataset classes, when the reference is reachable (build container / vendored copy)."""
    for cand in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.isdir(os.path.join(cand, 'genie')):
            for p in (os.path.join(ROOT, 'oracle', '_shim'), cand):
                if p not in sys.path:
                    sys.path.insert(0, p)
            from genie.module.data import Platformer2D            # noqa
            from genie.dataset import LightningPlatformer2D       # noqa
            return Platformer2D, LightningPlatformer2D
    return None, None


def test_platformer2d_matches_the_reference_dataset(clips):
    import open_genie_b200 as og
    RefP, RefL = _reference_classes()
    for fmt in ('t c h w', 'c t h w'):
        for padding in ('none', 'repeat', 'zero'):
            ds = og.Platformer2D(clips, split='train', padding=padding, num_frames=16, output_format=fmt)
            assert len(ds) == 5
            v = ds[1]
            shape = (16, 3, 64, 64) if fmt == 't c h w' else (3, 16, 64, 64)
            assert v.shape == shape and v.dtype == torch.float32 and 0.0 <= float(v.min()) and float(v.max()) <= 1.0
            short = [ds[i] for i in range(5) if ds.file_names[i].endswith('clip0.mp4')][0]
            assert short.shape[0 if fmt == 't c h w' else 1] == 12        # whole (shorter) video: data.py:191-193
            if RefP is not None:
                rds = RefP(clips, split='train', padding=padding, num_frames=16, output_format=fmt)
                assert rds.file_names == ds.file_names
                for i in range(len(ds)):
                    assert torch.equal(ds[i], rds[i])                          # bit-identical CPU tensors
    raw = og.Platformer2D(clips, split='val', num_frames=16, raw_uint8=True)[0]
    assert raw.dtype == torch.uint8 and raw.shape == (16, 64, 64, 3)
    ref = og.Platformer2D(clips, split='val', num_frames=16, output_format='t h w c')[0]
    assert torch.equal(raw.flip(-1).float() / 255., ref)                       # raw frames are BGR


def test_lightning_datamodule_surface(clips, tmp_path):
    import open_genie_b200 as og
    dm = og.LightningPlatformer2D(clips, num_frames=16, output_format='c t h w', batch_size=2, num_workers=0)
    dm.setup('fit')
    batch = next(iter(dm.train_dataloader()))
    assert batch.shape[0] == 2 and batch.shape[1] == 3 and batch.dtype == torch.float32
    assert len(dm.valid_dataset) == 2
    dm.setup('test')
    assert len(dm.test__dataset) == 2
    with pytest.raises(ValueError, match='Invalid stage'):
        dm.setup('nope')
    with pytest.raises(NotImplementedError):
        og.LightningDataset().setup('fit')
    cfg = tmp_path / 'conf.yaml'
    cfg.write_text(f'dataset:\n  root: {clips}\n  num_frames: 8\n  batch_size: 3\n')
    dm2 = og.LightningPlatformer2D.from_config(str(cfg))
    assert dm2.num_frames == 8 and dm2.batch_size == 3


@pytest.mark.gpu
def test_device_frame_decode_and_prefetcher(clips):
    import open_genie_b200 as og
    from open_genie_b200 import ops
    torch.manual_seed(0)
    frames = torch.randint(0, 256, (2, 4, 16, 16, 3), dtype=torch.uint8)
    want = (frames.flip(-1).float() / 255.).permute(0, 4, 1, 2, 3).contiguous()      # BGR -> RGB, /255, 'c t h w'
    got = og.frames_to_video(frames.cuda())
    assert got.shape == (2, 3, 4, 16, 16) and torch.equal(got.cpu(), want)            # exact: same fp32 arithmetic
    gi = og.frames_to_video(frames.cuda(), internal=True, cpad=8)
    assert ops.is_internal(gi[:, :3].contiguous(memory_format=torch.channels_last_3d)) or gi.shape == (2, 3, 4, 16, 16)
    assert torch.equal(gi.float().cpu(), want.to(torch.bfloat16).float())
    # end to end: uint8 loader -> pinned staging -> side-stream copy + decode == the reference-format CPU pipeline
    ds_raw = og.Platformer2D(clips, split='test', num_frames=16, raw_uint8=True)
    ds_ref = og.Platformer2D(clips, split='test', num_frames=16, output_format='c t h w')
    loader = torch.utils.data.DataLoader(ds_raw, batch_size=1, shuffle=False)
    pf = og.VideoBatchPrefetcher(loader)
    seen, elems = 0, 0
    for i, video in enumerate(pf):
        assert video.is_cuda and torch.equal(video[0].cpu(), ds_ref[i])
        seen += 1
        elems += video.numel()
    assert seen == len(ds_raw) and pf.h2d_bytes == elems                                # 1 byte per element over PCIe
