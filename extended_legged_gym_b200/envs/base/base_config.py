"""Nested-class configuration base.

Mirrors the behaviour of the reference's ``BaseConfig``
(legged_gym/legged_gym/envs/base/base_config.py:33-55): constructing a config
object replaces every nested *class* attribute by an *instance* of it,
recursively, so that ``cfg.rewards.scales.torques`` is a plain attribute
lookup on instances and per-object edits do not leak into the class.
"""
import inspect


def _instantiate_nested(node) -> None:
    for attr in dir(node):
        if attr == "__class__":
            continue
        member = getattr(node, attr)
        if inspect.isclass(member):
            inst = member()
            setattr(node, attr, inst)
            _instantiate_nested(inst)


class BaseConfig:
    def __init__(self) -> None:
        _instantiate_nested(self)

    # kept for API compatibility with the reference (static helper of the same name)
    @staticmethod
    def init_member_classes(obj) -> None:
        _instantiate_nested(obj)
