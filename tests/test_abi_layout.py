"""The ctypes mirrors of the C-ABI argument structs (openrl_b200/lib.py) have exactly the layout
gcc gives the structs declared in include/openrl_b200.h (size and every field offset)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def _c_layout(struct, fields, tmp_path):
    src = tmp_path / f"{struct}.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "openrl_b200.h"', "int main(void) {",
             f'  printf("sizeof %zu\\n", sizeof({struct}));']
    for f in fields:
        lines.append(f'  printf("{f} %zu\\n", offsetof({struct}, {f}));')
    lines += ["  return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / f"{struct}.bin"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    return dict((k, int(v)) for k, v in (ln.split() for ln in out.strip().splitlines()))


def test_struct_layouts_match_header(tmp_path):
    from openrl_b200 import lib

    for name, cls in (("OrlRolloutArgs", lib.OrlRolloutArgs), ("OrlPpoArgs", lib.OrlPpoArgs),
                      ("OrlRnnArgs", lib.OrlRnnArgs), ("OrlPeerArgs", lib.OrlPeerArgs),
                      ("OrlSelfPlayArgs", lib.OrlSelfPlayArgs)):
        fields = [f[0] for f in cls._fields_]
        c = _c_layout(name, fields, tmp_path)
        assert c["sizeof"] == ctypes.sizeof(cls), name
        for f in fields:
            assert c[f] == getattr(cls, f).offset, (name, f)


def test_header_flag_values_match_python():
    from openrl_b200 import lib

    text = open(os.path.join(ROOT, "include", "openrl_b200.h")).read()
    defs = dict((k, int(v)) for k, v in re.findall(r"#define (ORL_[A-Z0-9_]+) (\d+)\b", text))
    pairs = {"ORL_ENV_NONE": lib.ENV_NONE, "ORL_ENV_CARTPOLE": lib.ENV_CARTPOLE, "ORL_ENV_GRIDWORLD": lib.ENV_GRIDWORLD,
             "ORL_ENV_MPE_SPREAD": lib.ENV_MPE_SPREAD, "ORL_HEAD_CATEGORICAL": lib.HEAD_CATEGORICAL,
             "ORL_HEAD_GAUSSIAN": lib.HEAD_GAUSSIAN, "ORL_GAE_USE_GAE": lib.GAE_USE_GAE,
             "ORL_GAE_PROPER_TIME_LIMITS": lib.GAE_PROPER_TIME_LIMITS, "ORL_GAE_DENORM": lib.GAE_DENORM,
             "ORL_PPO_HUBER": lib.PPO_HUBER, "ORL_PPO_CLIP_VALUE": lib.PPO_CLIP_VALUE,
             "ORL_PPO_VALUE_ACTIVE_MASKS": lib.PPO_VALUE_ACTIVE_MASKS, "ORL_PPO_POLICY_ACTIVE_MASKS": lib.PPO_POLICY_ACTIVE_MASKS,
             "ORL_PPO_VALUENORM": lib.PPO_VALUENORM, "ORL_PPO_ADV_NORMALIZE": lib.PPO_ADV_NORMALIZE,
             "ORL_PPO_MAX_GRAD_NORM": lib.PPO_MAX_GRAD_NORM, "ORL_PPO_TENSORCORE": lib.PPO_TENSORCORE, "ORL_PPO_A2C": lib.PPO_A2C,
             "ORL_PPO_DUAL_CLIP": lib.PPO_DUAL_CLIP, "ORL_PEER_MAX_WORLD": lib.PEER_MAX_WORLD,
             "ORL_ENV_GRIDWORLD_2P": lib.ENV_GRIDWORLD_2P, "ORL_SP_RANDOM": lib.SP_RANDOM, "ORL_SP_LAST": lib.SP_LAST}
    for k, v in pairs.items():
        assert defs[k] == v, k
