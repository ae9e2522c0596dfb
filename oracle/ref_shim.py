"""Import the UNMODIFIED reference modules (read-only, /root/reference) with the minimum shims needed on
Python 3.12 / NumPy 2 without pybullet / parl / rlschool.

Only usable in the build container (the GPU box has no /root/reference): used by
tests/golden/make_golden.py to generate the committed fixtures and by the `ref`-marked tests that
re-check the oracle against the live reference when it is present.
"""
import ast
import collections
import collections.abc
import os
import sys
import types

REF = "/root/reference/QuadrupedalRobots/ETGRL"


def available():
    return os.path.isdir(REF)


_loaded = {}


def load():
    """Returns a namespace with the reference modules/functions that are importable here."""
    if _loaded:
        return _loaded["ns"]
    import numpy as np
    import torch

    collections.Sequence = collections.abc.Sequence  # minitaur.py:199, laikago_motor.py:62
    pb = types.ModuleType("pybullet")
    pb.getQuaternionFromEuler = lambda e: (0, 0, 0, 1)  # laikago_constants.py:35 runs it at import
    sys.modules.setdefault("pybullet", pb)
    parl = types.ModuleType("parl")
    parl.Model = torch.nn.Module
    parl.Algorithm = type("Algorithm", (), {})
    parl.Agent = type("Agent", (), {"__init__": lambda self, alg: setattr(self, "alg", alg)})
    sys.modules.setdefault("parl", parl)
    for m in ("motion_imitation", "motion_imitation.robots"):
        sys.modules.setdefault(m, types.ModuleType(m))
    for p in (REF + "/deployment", REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from robots import robot_config

        sys.modules["motion_imitation.robots"].robot_config = robot_config
        from robots import laikago_motor, a1, action_filter, minitaur
        from model.mujoco_model import MujocoModel
        from model.mujoco_agent import MujocoAgent
        from alg.sac import SAC
        from alg import es

    # train.py helpers: exec only the pure FunctionDefs (module import needs rlschool)
    src = open(REF + "/train.py").read()
    tree = ast.parse(src)
    helpers = {"np": np, "copy": __import__("copy").copy}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("LS_sol", "Opt_with_points", "param2dynamic_dict"):
            exec(compile(ast.Module([node], []), REF + "/train.py", "exec"), helpers)
    ns = types.SimpleNamespace(
        robot_config=robot_config, laikago_motor=laikago_motor, a1=a1, action_filter=action_filter, minitaur=minitaur,
        MujocoModel=MujocoModel, MujocoAgent=MujocoAgent, SAC=SAC, es=es,
        LS_sol=helpers["LS_sol"], Opt_with_points=helpers["Opt_with_points"], param2dynamic_dict=helpers["param2dynamic_dict"],
        REF=REF,
    )
    _loaded["ns"] = ns
    return ns
