"""Pre-tokenised video-index format (nuwa_pytorch_b200/data.py) against bytes written by the UNMODIFIED reference
(train_nuwa.py:56-80 through oracle/make_golden_data.py): the product writer must produce the identical file, for any
batching, and the reader / collate must return what the reference's do."""
import json
import os

import numpy as np
import pytest
import torch

from nuwa_pytorch_b200.data import VideoIndicesDataset, convert_video_tensor_dataset_to_indices, pad_collate_fn
from tests.helpers_data import StubVAE, StubVideos

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("batch_videos", [1, 2, 8])
def test_writer_bytes_match_reference(tmp_path, batch_videos):
    vae = StubVAE(image_size=64, num_layers=4, codebook=97)
    videos = StubVideos(n=5, frames=3, channels=3, size=64, seed=7)
    path = str(tmp_path / "idx.bin")
    shape = convert_video_tensor_dataset_to_indices(vae=vae, raw_video_dataset=videos, num_frames=3, path=path,
                                                    batch_videos=batch_videos)
    meta = json.load(open(os.path.join(G, "video_indices_small.json")))
    assert list(shape) == meta["shape"]
    assert open(path, "rb").read() == open(os.path.join(G, "video_indices_small.bin"), "rb").read()
    assert vae.calls == [min(batch_videos, 5 - s) for s in range(0, 5, batch_videos)]  # videos really were batched


def test_reader_and_collate_match_reference():
    vae = StubVAE(image_size=64, num_layers=4, codebook=97)
    meta = json.load(open(os.path.join(G, "video_indices_small.json")))
    ds = VideoIndicesDataset(videos_memmap_path=os.path.join(G, "video_indices_small.bin"),
                             text_memmap_path=os.path.join(G, "video_indices_small_text.bin"), vae=vae, num_videos=5, num_frames=3)
    assert len(ds) == 5
    text, video = ds[3]
    assert video.dtype == torch.int64 and video.shape == (48,)
    assert int(video.sum()) == meta["item3_video_sum"] and video[:8].tolist() == meta["item3_first8"]
    # default encoder: digit + 1 (id 0 is NUWA's pad id and would be masked out of the conditioning -- ADVICE r1);
    # the reference's BPE tokenizer is out of scope and can be passed as text_encode
    assert text.tolist() == [7, 8]
    assert all(int(t) != 0 for i in range(len(ds)) for t in ds[i][0])
    bad = VideoIndicesDataset(videos_memmap_path=os.path.join(G, "video_indices_small.bin"),
                              text_memmap_path=os.path.join(G, "video_indices_small_text.bin"), vae=vae, num_videos=5, num_frames=3,
                              text_encode=lambda lab: [int(d) for d in lab])
    with pytest.raises(AssertionError, match='padding'):
        [bad[i] for i in range(len(bad))]   # some label holds a 0 digit -> id 0 -> rejected
    ds2 = VideoIndicesDataset(videos_memmap_path=os.path.join(G, "video_indices_small.bin"),
                              text_memmap_path=os.path.join(G, "video_indices_small_text.bin"), vae=vae, num_videos=5, num_frames=3,
                              text_encode=lambda lab: [49406] + [100 + d for d in lab] + [49407])
    assert ds2[0][0].tolist() == [49406, 100, 101, 49407]
    t, v = pad_collate_fn([(torch.tensor([1, 2, 3]), ds[3][1]), (torch.tensor([4]), ds[0][1])])
    assert t.tolist() == meta["collate_text"] and list(v.shape) == meta["collate_video_shape"]


def test_writer_rejects_empty_and_mismatched_rows(tmp_path):
    vae = StubVAE(image_size=64, num_layers=4, codebook=97)
    with pytest.raises(AssertionError):
        convert_video_tensor_dataset_to_indices(vae=vae, raw_video_dataset=[], num_frames=3, path=str(tmp_path / "a.bin"))
    videos = StubVideos(n=2, frames=3, channels=3, size=64, seed=1)
    with pytest.raises(AssertionError):  # file rows sized for 2 frames, videos have 3
        convert_video_tensor_dataset_to_indices(vae=vae, raw_video_dataset=videos, num_frames=2, path=str(tmp_path / "b.bin"))
    assert np.memmap(str(tmp_path / "b.bin"), dtype=np.int64, mode="r").shape[0] == 2 * 2 * 16
