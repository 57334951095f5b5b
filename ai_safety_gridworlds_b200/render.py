"""RGB rendering of the boards (the `RGB` entry of the reference's observation distiller).

`ObservationToArrayWithRGBEx.__call__` (environments/shared/observation_distiller_ex.py:147-189) paints every board
character with the game's colour table (GAME_BG_COLOURS updated with the shared tables; pycolab's 0..999 scale) through
pycolab's `ObservationToArray` (pycolab/rendering.py:491-549) and scales it: `(RGB / 999.0 * 255.0).astype(np.uint8)`,
shape [3, H, W].  Here the same table look-up runs as one kernel over a whole batch of boards (`gw_render_rgb`,
include/gwsim.h); the per-game colour constants were read out of the running reference by oracle/dump_colours.py into
envs/colours.json.  No CPU fallback.
"""
import ctypes as C
import json
import os

import numpy as np

from . import _abi

_COLOURS = None


def colour_mapping(env_name):
    """{character: (r, g, b)} on pycolab's 0..999 scale for a factory name (experiment overlays included)."""
    global _COLOURS
    if _COLOURS is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "envs", "colours.json")) as f:
            _COLOURS = json.load(f)
    try:
        return {ch: tuple(v) for ch, v in _COLOURS[env_name.lower()].items()}
    except KeyError:
        raise NotImplementedError("no colour table for environment %r" % (env_name,))


def rgb_lut(env_name):
    """uint8 [256, 3]: the reference's scaling of the colour table, character code -> (R, G, B); characters the game has no
    colour for map to black."""
    lut = np.zeros((256, 3), np.float64)
    for ch, rgb in colour_mapping(env_name).items():
        if len(ch) == 1 and ord(ch) < 256:
            lut[ord(ch)] = rgb
    return (lut / 999.0 * 255.0).astype(np.uint8)


def render_rgb(board, lut):
    """board: uint8 CUDA tensor [N, H, W] (or [N, cells]; the last dimensions dense, the environment stride arbitrary) of
    ASCII codes; lut: uint8 CUDA tensor [256, 3].  Returns uint8 [N, 3, H, W] (or [N, 3, cells])."""
    import torch
    if board.dtype != torch.uint8 or not board.is_cuda:
        raise ValueError("board must be a uint8 CUDA tensor")
    n = board.shape[0]
    inner = tuple(board.shape[1:])
    cells = int(np.prod(inner))
    flat = board.reshape(n, cells) if board[0].is_contiguous() else board.contiguous().reshape(n, cells)
    if flat.stride(1) != 1:
        flat = flat.contiguous()
    pitch = flat.stride(0) if n > 1 else cells
    lut = lut.to(device=board.device, dtype=torch.uint8).contiguous()
    out = torch.empty((n, 3) + inner, dtype=torch.uint8, device=board.device)
    stream = torch.cuda.current_stream(board.device).cuda_stream
    _abi.check(_abi.load().gw_render_rgb(C.c_void_p(flat.data_ptr()), n, cells, int(pitch), C.c_void_p(lut.data_ptr()),
                                         C.c_void_p(out.data_ptr()), board.device.index or 0, C.c_void_p(stream)))
    return out
