"""TEST INFRASTRUCTURE ONLY (build container): run the UNMODIFIED reference writer / reader of the pre-tokenised
video-index format (train_nuwa.py:56-80,120-147) on a deterministic stub VAE and store the bytes it produces as
tests/golden/video_indices_small.* -- the pin for nuwa_pytorch_b200/data.py."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402
from tests.helpers_data import StubVAE, StubVideos  # noqa: E402

import_reference()
import nuwa_pytorch.train_nuwa as TN  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
vae = StubVAE(image_size=64, num_layers=4, codebook=97)
videos = StubVideos(n=5, frames=3, channels=3, size=64, seed=7)
path = os.path.join(G, "video_indices_small.bin")
TN.convert_video_tensor_dataset_to_indices(vae=vae, raw_video_dataset=videos, num_frames=3, path=path)
labels = np.memmap(os.path.join(G, "video_indices_small_text.bin"), mode='w+', dtype=np.uint8, shape=(5, 2))
labels[:] = np.arange(10, dtype=np.uint8).reshape(5, 2)
labels.flush()
ds = TN.VideoIndicesDataset(videos_memmap_path=path, text_memmap_path=os.path.join(G, "video_indices_small_text.bin"), vae=vae,
                            num_videos=5, num_frames=3)
# the reference reader tokenises the label with its BPE tokenizer; record the raw pieces it reads instead
item_video = ds[3][1]
batch = TN.pad_collate_fn([(torch.tensor([1, 2, 3]), item_video), (torch.tensor([4]), ds[0][1])])
json.dump(dict(shape=[5, 3 * 4 * 4], item3_video_sum=int(item_video.sum()), item3_first8=item_video[:8].tolist(),
               collate_text=batch[0].tolist(), collate_video_shape=list(batch[1].shape),
               sha_note="bytes of video_indices_small.bin are the golden"),
          open(os.path.join(G, "video_indices_small.json"), "w"), indent=1)
print("wrote", path, os.path.getsize(path), "bytes")
