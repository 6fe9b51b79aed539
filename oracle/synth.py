"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights.

Golden fixtures cannot carry multi-MB checkpoints, so every fixture stores only a *manifest*
(key, shape, dtype) plus a seed; `synth_state_dict` regenerates identical fp32 tensors on any machine
with the same torch build (CPU generator).  The same tensors were loaded into the UNMODIFIED reference when
the fixture was produced (oracle/make_golden.py)."""
import math
import zlib

import torch


def _gen(key, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def synth_tensor(key, shape, seed):
    """Value policy (chosen so activations stay O(1) through deep stacks):
    norm gains ~ 1+0.1n, biases ~ 0.1n, linear/conv weights ~ n/sqrt(fan_in), embeddings / positional /
    null-kv / bos ~ n, codebook = l2-normalised n, VQGanAttention.scale = log(0.01)+0.1n."""
    g = _gen(key, seed)
    shape = tuple(shape)
    r = lambda: torch.randn(shape, generator=g)
    leaf = key.split('.')[-1]
    if leaf == 'cluster_size':
        return torch.zeros(shape)
    if leaf == 'initted':
        return torch.ones(shape)
    if key.endswith('_codebook.embed'):
        return torch.nn.functional.normalize(r(), dim=-1)
    if leaf == 'scale':
        return math.log(0.01) + 0.1 * r()
    if leaf == 'g':
        return 1 + 0.1 * r()
    if leaf == 'b':
        return 0.1 * r()
    if leaf == 'bias':
        return 0.1 * r()
    if leaf == 'weight':
        if len(shape) == 1:
            return 1 + 0.1 * r()
        if key.endswith('embed.weight'):
            return r()
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return r() / math.sqrt(fan_in)
    return r()  # video_bos, axial*, null_k, null_v, ...


def is_synth_key(key, tensor):
    """float tensors are synthesised; bool masks / inv_freq keep the values the module computed itself."""
    return tensor.dtype == torch.float32 and not key.endswith('inv_freq')


def manifest_of(state_dict):
    """One entry per distinct storage: ReversibleTransformer exposes every tensor twice (layers.* and
    net.blocks.*, nuwa_pytorch.py:1286 / reversible.py:130); only the first alias is synthesised."""
    seen, out = set(), []
    for k, v in state_dict.items():
        if not is_synth_key(k, v):
            continue
        key = (v.data_ptr(), tuple(v.shape))
        if v.numel() > 0 and key in seen:
            continue
        seen.add(key)
        out.append((k, tuple(v.shape)))
    return out


def synth_state_dict(manifest, seed):
    return {k: synth_tensor(k, shape, seed) for k, shape in manifest}
