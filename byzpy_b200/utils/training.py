"""``train_with_progress``: tqdm progress bar over ``ps.round()`` with periodic evaluation
(reference utils/training.py:7-34)."""
from __future__ import annotations

import inspect
from typing import Any, Awaitable, Callable, Dict, Optional, Union

EvalFn = Callable[[], Union[Dict[str, Any], Awaitable[Dict[str, Any]], None]]


async def train_with_progress(ps, rounds: int, eval_callback: Optional[EvalFn] = None,
                              eval_interval: int = 50, *, desc: str = "training") -> None:
    try:
        from tqdm import tqdm
    except Exception:  # pragma: no cover
        tqdm = None
    bar = tqdm(range(1, rounds + 1), desc=desc) if tqdm is not None else None
    for r in (bar if bar is not None else range(1, rounds + 1)):
        await ps.round()
        if eval_callback is not None and eval_interval > 0 and r % eval_interval == 0:
            metrics = eval_callback()
            if inspect.isawaitable(metrics):
                metrics = await metrics
            if bar is not None and isinstance(metrics, dict):
                bar.set_postfix(metrics)
    if bar is not None:
        bar.close()


__all__ = ["train_with_progress"]
