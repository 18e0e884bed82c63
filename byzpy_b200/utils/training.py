"""``train_with_progress``: tqdm progress bar over ``ps.round()`` with periodic evaluation
(reference utils/training.py:7-34)."""
from __future__ import annotations

import inspect
from typing import Any, Awaitable, Callable, Dict, List, Optional, Union

EvalFn = Callable[..., Union[Dict[str, Any], Awaitable[Dict[str, Any]], None]]


def _wants_round(fn: Callable[..., Any]) -> bool:
    """True when ``fn`` can take the round number as a positional argument."""
    try:
        params = list(inspect.signature(fn).parameters.values())
    except (TypeError, ValueError):
        return False
    return any(p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD, p.VAR_POSITIONAL) for p in params)


async def train_with_progress(ps, rounds: int, eval_callback: Optional[EvalFn] = None,
                              eval_interval: int = 1, *, desc: str = "training") -> List[Dict[str, Any]]:
    """Run ``rounds`` rounds of ``ps`` under a progress bar.  Every ``eval_interval`` rounds
    ``eval_callback(round_number)`` (sync or async; a zero-argument callable is accepted too) is evaluated,
    its metrics dict shown on the bar and appended to the returned history as
    ``{"round": r, "metrics": metrics}``.  Without a callback the history is empty."""
    try:
        from tqdm import tqdm
    except Exception:  # pragma: no cover
        tqdm = None
    history: List[Dict[str, Any]] = []
    takes_round = eval_callback is not None and _wants_round(eval_callback)
    bar = tqdm(range(1, rounds + 1), desc=desc) if tqdm is not None else None
    try:
        for r in (bar if bar is not None else range(1, rounds + 1)):
            await ps.round()
            if eval_callback is not None and eval_interval > 0 and r % eval_interval == 0:
                metrics = eval_callback(r) if takes_round else eval_callback()
                if inspect.isawaitable(metrics):
                    metrics = await metrics
                history.append({"round": r, "metrics": metrics})
                if bar is not None and isinstance(metrics, dict):
                    bar.set_postfix(metrics)
    finally:
        if bar is not None:
            bar.close()
    return history


__all__ = ["train_with_progress"]
