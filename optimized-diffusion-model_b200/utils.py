"""Checkpoint ingestion -- drop-in for the inference-side functions of the reference's `utils` module
(/root/reference/Reflected-Diffusion/utils.py:48-89).  Reference checkpoints are `torch.save`d dicts
`{step, model, optimizer, ema{decay,num_updates,shadow_params}, scaler, config}`; the B200 NCSNpp keeps the
reference's state_dict keys, so `model` and `ema` load unchanged.  Hydra/logging helpers are out of scope.
"""
import logging
import os

import torch


def makedirs(dirname):
    os.makedirs(dirname, exist_ok=True)


def _unwrap(model):
    return model.module if hasattr(model, 'module') else model


def restore_checkpoint(ckpt_dir, state, device, ddp=True):
    """Fill `state` (keys: model, ema, step; optional optimizer, scaler) from a reference checkpoint.
    A missing file logs a warning and returns `state` untouched, like the reference (utils.py:48-54).
    The optimizer / scaler entries are restored only when `state` carries live objects for them (the
    inference path has none)."""
    if not os.path.exists(ckpt_dir):
        makedirs(os.path.dirname(ckpt_dir))
        logging.warning(f"No checkpoint found at {ckpt_dir}. Returned the same state as input")
        return state
    loaded = torch.load(ckpt_dir, map_location=device, weights_only=False)
    if state.get('optimizer') is not None:
        state['optimizer'].load_state_dict(loaded['optimizer'])
    _unwrap(state['model']).load_state_dict(loaded['model'], strict=False)
    state['ema'].load_state_dict(loaded['ema'])
    state['step'] = loaded['step']
    if state.get('scaler') is not None:
        state['scaler'].load_state_dict(loaded['scaler'])
    return state


def load_denoising_model(ckpt_dir, model, device=torch.device('cpu')):
    """Load only the `model` entry (utils.py:69-74); a missing file raises ValueError."""
    if not os.path.exists(ckpt_dir):
        raise ValueError(f"No checkpoint found at {ckpt_dir}.")
    loaded = torch.load(ckpt_dir, map_location=device, weights_only=False)
    model.load_state_dict(loaded['model'], strict=False)
    return model


def save_checkpoint(ckpt_dir, state):
    """Write the reference's checkpoint layout (utils.py:77-86)."""
    model = _unwrap(state['model'])
    torch.save({
        'step': state['step'],
        'model': model.state_dict(),
        'optimizer': state['optimizer'].state_dict() if state.get('optimizer') is not None else None,
        'ema': state['ema'].state_dict() if 'ema' in state else None,
        'scaler': state['scaler'].state_dict() if state.get('scaler') is not None else None,
        'config': state.get('config'),
    }, ckpt_dir)
