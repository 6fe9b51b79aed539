"""Pre-tokenised video-index data format either side of the decoder hot path.

On-disk format (reference: train_nuwa.py:56-80 writer, :120-147 reader): a raw little-endian int64 array of shape
(num_videos, num_frames * fmap * fmap), one row per video, row = '(f h w)'-flattened VQ codebook indices;
fmap = vae.image_size // vae.num_layers ** 2 (the reference's own expression, SURVEY D5 -- identical to the real
feature-map side for the 4-layer VAEs of every BASELINE config).  Text labels live in a separate uint8 array
(num_videos, num_digits).  No header: shapes travel out of band, exactly as in the reference.

B200-first differences from the reference writer (same bytes on disk):
  * videos are encoded in batches of `batch_videos` (the reference runs one video per `get_video_indices` call:
    at 10 frames that is a 10-frame VAE batch -- launch bound; 64-frame batches run the conv stack at 0.87 of the
    tensor peak);
  * frames are staged through one pinned host buffer and copied asynchronously; indices come back through a pinned
    buffer and are written to the memmap row by row.
The VAE passed in is the product `nuwa_pytorch_b200.VQGanVAE` on a CUDA device (no CPU path); any object with
`image_size`, `num_layers`, `parameters()` and `get_video_indices(video) -> (b, f, h, w) int64` works (the CPU tests
use a stub so that the format logic is covered without a GPU).
"""
import os

import numpy as np
import torch
from torch.utils.data import Dataset

from .parallel import rank_slice


def _fmap_size(vae):
    return vae.image_size // (vae.num_layers ** 2)  # train_nuwa.py:67 (D5 kept)


def convert_video_tensor_dataset_to_indices(*, vae, raw_video_dataset, num_frames, path, batch_videos=8, rank=0, world_size=1,
                                            barrier=None):
    """Same contract as train_nuwa.py:56-80: `raw_video_dataset[i] -> (text, video (f, c, h, w) float)`; writes the
    int64 memmap at `path` and returns its shape.

    Multi-GPU (one process per GPU): the videos are independent units, so rank r of `world_size` encodes the contiguous
    slice `parallel.rank_slice(n, r, world)` with its own VAE replica and writes those rows of the SAME file with
    positioned writes on its own file handle (safe across hosts on shared filesystems) -- no data-path collective; `barrier` (e.g. `torch.distributed.barrier`) is called once after rank 0 has created the file and once
    after every rank has flushed its rows."""
    try:
        device = next(vae.parameters()).device
    except StopIteration:
        device = torch.device('cpu')
    num_videos = len(raw_video_dataset)
    assert num_videos > 0, 'there must be at least 1 video'
    assert 0 <= rank < world_size and (world_size == 1 or barrier is not None), 'multi-rank conversion needs a barrier callable'
    fmap = _fmap_size(vae)
    shape = (num_videos, num_frames * fmap * fmap)
    if rank == 0:
        out = np.memmap(path, mode='w+', dtype=np.int64, shape=shape)  # creates the file at full size
        out.flush()
        del out
    if world_size > 1:
        barrier()  # the file exists at full size before any other rank opens it
    lo, hi = rank_slice(num_videos, rank, world_size)  # the same owner map as every other sharded path (parallel.py)
    row_bytes = shape[1] * 8
    pinned = device.type == 'cuda'
    stage = None
    # Each rank writes ITS rows through its own file handle with positioned writes.  (A shared mmap whose row
    # boundaries are not page aligned is only coherent between processes of one host; pwrite of disjoint byte ranges
    # is safe on any POSIX filesystem, NFS / Lustre included.)
    fd = os.open(path, os.O_WRONLY)
    try:
        for start in range(lo, hi, batch_videos):
            vids = [raw_video_dataset[i][1] for i in range(start, min(start + batch_videos, hi))]
            batch = torch.stack(vids)
            if pinned:
                if stage is None or stage.shape[1:] != batch.shape[1:] or stage.dtype != batch.dtype:
                    stage = torch.empty((batch_videos,) + tuple(batch.shape[1:]), dtype=batch.dtype).pin_memory()
                stage[:batch.shape[0]].copy_(batch)
                batch = stage[:batch.shape[0]].to(device, non_blocking=True)
            indices = vae.get_video_indices(batch)  # (b, f, h, w) int64
            assert indices.shape[1] * indices.shape[2] * indices.shape[3] == shape[1], \
                f'VAE produced {tuple(indices.shape[1:])} indices per video, the file row holds {shape[1]}'
            flat = indices.reshape(indices.shape[0], -1).to('cpu', torch.int64).contiguous()
            buf = flat.numpy().astype('<i8', copy=False).tobytes()
            off, view = start * row_bytes, memoryview(buf)
            while view:  # pwrite may be partial
                n = os.pwrite(fd, view, off)
                off, view = off + n, view[n:]
        os.fsync(fd)
    finally:
        os.close(fd)
    if world_size > 1:
        barrier()  # every rank's rows are on disk
    return shape


def digits_plus_one(label):
    """Default text encoder of VideoIndicesDataset: label digit d -> token id d + 1.

    Id 0 is the PAD id that NUWA.forward / generate mask out (`text != 0`, nuwa_pytorch.py:1936,1862), so a label digit
    0 must not map to it (it would silently drop the digit from the conditioning; a label like [0, 0] would become
    unconditional).  The reference tokenises ' '.join(digits) with its BPE tokenizer (train_nuwa.py:143; needs `ftfy`
    and the vocabulary file, out of scope), whose ids are never 0 either -- pass that callable as `text_encode` to
    reproduce it exactly."""
    return [int(x) + 1 for x in label]


class VideoIndicesDataset(Dataset):
    """Reader of the format above (train_nuwa.py:120-147): item = (text ids int64 (n,), video indices int64 (F*fmap^2,))."""

    def __init__(self, *, videos_memmap_path, text_memmap_path, vae, num_videos, num_frames, num_digits=2,
                 text_encode=digits_plus_one):
        self.num_videos = num_videos
        fmap = _fmap_size(vae)
        self.videos_memmap = np.memmap(videos_memmap_path, mode='r', dtype=np.int64, shape=(num_videos, num_frames * fmap ** 2))
        self.text_memmap = np.memmap(text_memmap_path, mode='r', dtype=np.uint8, shape=(num_videos, num_digits))
        self.text_encode = text_encode

    def __len__(self):
        return self.num_videos

    def __getitem__(self, idx):
        video = torch.from_numpy(self.videos_memmap[idx].copy()).long()
        label = self.text_memmap[idx].copy().tolist()
        ids = self.text_encode(label)
        assert all(int(t) != 0 for t in ids), 'text_encode produced id 0, which NUWA treats as padding (masked out)'
        text = torch.tensor(ids, dtype=torch.long)
        return text, video


def pad_collate_fn(batch):
    """train_nuwa.py:50-52: right-pad the texts with 0 (the pad id NUWA.forward masks, nuwa_pytorch.py:1936), stack videos."""
    texts, videos = zip(*batch)
    return torch.nn.utils.rnn.pad_sequence(texts, batch_first=True), torch.stack(videos)
