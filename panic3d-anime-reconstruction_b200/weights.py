"""Flat weight files for the generator - the load side of SURVEY.md 8f-4.

The reference stores a training snapshot as a pickle of live module objects (``legacy.load_network_pkl`` through
``torch_utils.persistence``, reference ``_train/eg3dc/util/eg3dc_v0.py:41-52``): un-pickling executes the pickled source of
every module, builds each parameter as its own small CPU tensor, then ``load_eg3dc_model`` re-instantiates
``TriPlaneGenerator(*G.init_args, **G.init_kwargs)`` and copies parameter by parameter
(``misc.copy_params_and_buffers(G, G_new, require_all=True)``); ``.to(device)`` afterwards issues one small H2D copy per
tensor (~300 for G).  A flat file keeps what that procedure actually needs:

    header   magic, JSON: tensor table (name, dtype, shape, offset, nbytes) + meta (init_args, init_kwargs,
             neural_rendering_resolution, rendering_kwargs - the attributes load_eg3dc_model carries over)
    payload  every tensor's bytes at a 256-byte aligned offset

``load_weights`` reads the payload with ONE read into (pinned) host memory and moves it with ONE host->device copy; the
returned tensors are views into that single device buffer.  ``export_generator`` / ``build_generator`` mirror the
reload step of ``load_eg3dc_model`` (same ``require_all`` semantics).  Values round-trip bit for bit.
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np
import torch

MAGIC = b'P3DW0001'
ALIGN = 256
_DTYPES = {'float32': torch.float32, 'float16': torch.float16, 'bfloat16': torch.bfloat16, 'float64': torch.float64,
           'int64': torch.int64, 'int32': torch.int32, 'int16': torch.int16, 'int8': torch.int8, 'uint8': torch.uint8, 'bool': torch.bool}


def _round_up(n, a=ALIGN):
    return (n + a - 1) // a * a


def _jsonable(obj):
    """EasyDict / tuples / numpy scalars -> plain JSON types (what init_kwargs and rendering_kwargs hold)."""
    if isinstance(obj, dict):
        return {str(k): _jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_jsonable(v) for v in obj]
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return float(obj)
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return obj
    raise TypeError(f'panic3d_b200.weights: cannot store {type(obj).__name__} in the header')


def save_weights(state, path, meta=None):
    """``state``: an ``nn.Module`` (parameters + buffers, like ``misc.named_params_and_buffers``) or a name -> tensor dict."""
    if isinstance(state, torch.nn.Module):
        state = dict(list(state.named_parameters()) + list(state.named_buffers()))
    table, off = [], 0
    for name, t in state.items():
        dt = str(t.dtype).replace('torch.', '')
        if dt not in _DTYPES:
            raise TypeError(f'panic3d_b200.weights: {name} has unsupported dtype {t.dtype}')
        nbytes = t.numel() * t.element_size()
        table.append({'name': name, 'dtype': dt, 'shape': list(t.shape), 'offset': off, 'nbytes': nbytes})
        off = _round_up(off + nbytes)
    header = json.dumps({'tensors': table, 'meta': _jsonable(meta or {}), 'payload_bytes': off}).encode()
    start = _round_up(len(MAGIC) + 8 + len(header))
    tmp = os.fspath(path) + '.tmp'
    with open(tmp, 'wb') as f:
        f.write(MAGIC + struct.pack('<Q', len(header)) + header)
        f.write(b'\0' * (start - f.tell()))
        for entry, t in zip(table, state.values()):
            f.seek(start + entry['offset'])
            f.write(t.detach().cpu().contiguous().reshape(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b'')
        f.truncate(start + off)
    os.replace(tmp, path)
    return start + off


def read_header(path):
    with open(path, 'rb') as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise RuntimeError(f'panic3d_b200.weights: {path} is not a P3DW file')
        (n,) = struct.unpack('<Q', f.read(8))
        header = json.loads(f.read(n))
    return header, _round_up(len(MAGIC) + 8 + n)


def load_weights(path, device='cuda'):
    """-> (name -> tensor dict, meta).  One file read, one host->device copy; the tensors are views of one buffer."""
    header, start = read_header(path)
    total = header['payload_bytes']
    device = torch.device(device)
    pin = device.type == 'cuda'
    host = torch.empty(total, dtype=torch.uint8, pin_memory=pin)
    with open(path, 'rb') as f:
        f.seek(start)
        got = f.readinto(host.numpy()) if total else 0
    if got != total:
        raise RuntimeError(f'panic3d_b200.weights: {path} is truncated ({got} of {total} payload bytes)')
    buf = host.to(device, non_blocking=pin) if pin else host
    out = {}
    for e in header['tensors']:
        raw = buf[e['offset']:e['offset'] + e['nbytes']]
        out[e['name']] = raw.view(_DTYPES[e['dtype']]).view(e['shape']) if e['nbytes'] else torch.empty(e['shape'], dtype=_DTYPES[e['dtype']], device=device)
    if pin:
        torch.cuda.current_stream(device).synchronize()                       # the pinned staging dies with this frame
    return out, header['meta']


def load_into(module, path, device=None, require_all=True):
    """``misc.copy_params_and_buffers(src, module, require_all)`` (reference torch_utils/misc.py) from a flat file."""
    device = device or next(iter(module.parameters())).device
    tensors, meta = load_weights(path, device)
    with torch.no_grad():
        for name, t in list(module.named_parameters()) + list(module.named_buffers()):
            if name in tensors:
                t.copy_(tensors[name].to(t.dtype).view_as(t) if tensors[name].shape != t.shape else tensors[name])
            elif require_all:
                raise RuntimeError(f'panic3d_b200.weights: {path} has no tensor "{name}" (require_all)')
    return meta


def export_generator(G, path):
    """Everything ``load_eg3dc_model`` (eg3dc_v0.py:41-52) carries from the pickled ``G_ema`` into the rebuilt generator."""
    meta = {'init_args': list(getattr(G, 'init_args', ())), 'init_kwargs': dict(getattr(G, 'init_kwargs', {})),
            'neural_rendering_resolution': getattr(G, 'neural_rendering_resolution', None),
            'rendering_kwargs': dict(getattr(G, 'rendering_kwargs', {}))}
    return save_weights(G, path, meta)


def build_generator(path, generator_class, device='cuda', force_sigmoid=False, depth_resolution=48 * 2, depth_resolution_importance=48 * 2):
    """The reload branch of ``load_eg3dc_model`` from a flat file: ``generator_class(*init_args, **init_kwargs)`` in eval mode
    without gradients, all parameters and buffers required, then the attributes the reference copies / overrides."""
    header, _ = read_header(path)
    meta = header['meta']
    G = generator_class(*meta['init_args'], **meta['init_kwargs']).eval().requires_grad_(False).to(device)
    load_into(G, path, device=device, require_all=True)
    G.neural_rendering_resolution = meta['neural_rendering_resolution']
    G.rendering_kwargs = dict(meta['rendering_kwargs'])
    if force_sigmoid:
        G.set_force_sigmoid(True)
    G.rendering_kwargs['depth_resolution'] = depth_resolution
    G.rendering_kwargs['depth_resolution_importance'] = depth_resolution_importance
    return G
