"""Host-side helpers on the hot path's boundary.

``class_to_dict`` defines the reward-term ORDER: it walks ``dir(obj)``, which is
sorted, so the reward registry is alphabetical (reference:
legged_gym/legged_gym/utils/helpers.py:43-58).
"""
from typing import Any


def class_to_dict(obj: Any):
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for name in dir(obj):            # dir() is sorted -> alphabetical keys
        if name.startswith("_"):
            continue
        value = getattr(obj, name)
        out[name] = [class_to_dict(v) for v in value] if isinstance(value, list) else class_to_dict(value)
    return out


def update_class_from_dict(obj, d: dict) -> None:
    for key, val in d.items():
        attr = getattr(obj, key, None)
        if isinstance(val, dict) and attr is not None and hasattr(attr, "__dict__"):
            update_class_from_dict(attr, val)
        else:
            setattr(obj, key, val)
