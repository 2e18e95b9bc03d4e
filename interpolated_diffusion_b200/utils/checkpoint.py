"""On-disk checkpoint format of the reference (``src/utils/checkpoint.py:6-49``): one ``torch.save``d dict with the keys
``model`` and ``step`` and, when present, ``optimizer``, ``ema`` and ``meta``.  Anything with ``state_dict()`` /
``load_state_dict()`` can be stored in the optimizer / ema slots -- the reference's ``torch.optim.AdamW`` and ``EMA`` objects or
this package's ``FlatAdamW`` and ``FlatAdamW.ema_state`` (their state dicts have the same structure), so checkpoints move between
the two code bases in both directions.  Writes go through a temporary file and an atomic rename: a job killed mid-save never
leaves a truncated checkpoint behind."""
import os
from typing import Any, Dict, Optional, Tuple, Union

import torch

_OPTIONAL_SLOTS = ("optimizer", "ema")


def _payload(model: torch.nn.Module, step: int, slots: Dict[str, Any], meta: Optional[dict]) -> Dict[str, Any]:
    payload: Dict[str, Any] = {"model": model.state_dict(), "step": step}
    payload.update({name: obj.state_dict() for name, obj in slots.items() if obj is not None})
    if meta is not None:
        payload["meta"] = meta
    return payload


def save_checkpoint(path: str, model: torch.nn.Module, optimizer, step: int, ema: Optional[object] = None, meta: Optional[dict] = None,
                    *, save_optimizer: bool = True) -> None:
    slots = {"optimizer": optimizer if save_optimizer else None, "ema": ema}
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(_payload(model, step, slots, meta), tmp)
    os.replace(tmp, path)


def load_checkpoint(path: str, model: torch.nn.Module, optimizer=None, ema: Optional[object] = None, map_location: Optional[str] = None,
                    return_payload: bool = False) -> Union[int, Tuple[int, dict]]:
    """Restores ``model`` (and ``optimizer`` / ``ema`` when both the object and its slot exist); returns the stored step (0 if the
    file has none), with the whole payload if ``return_payload``."""
    payload = torch.load(path, map_location=map_location)
    model.load_state_dict(payload["model"])
    for name, obj in zip(_OPTIONAL_SLOTS, (optimizer, ema)):
        if obj is not None and name in payload:
            obj.load_state_dict(payload[name])
    step = payload.get("step", 0)
    return (step, payload) if return_payload else step
