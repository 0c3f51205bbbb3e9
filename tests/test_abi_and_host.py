"""CPU: the C-ABI library loads and exports every symbol include/gridworld_b200.h declares; host-side logic
(scenario tables, config struct layout, loud failure without a GPU).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import marl_responsible_nav_b200 as pkg
from marl_responsible_nav_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gridworld_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gww?_[a-z_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = N.load()
    syms = _declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in gridworld_b200.h but not exported"
    assert sorted(N.EXPORTS) == syms
    assert lib.gw_abi_version() == 1
    assert b"sm_100a" in lib.gw_build_info()


def test_config_struct_layout_matches_library():
    lib = N.load()
    cfg = N.GwConfig()
    assert lib.gw_default_config(C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(N.GwConfig)
    assert (cfg.height, cfg.width, cfg.n_agents, cfg.n_learners) == (10, 16, 4, 2)
    assert (cfg.apple_row[0], cfg.apple_col[0], cfg.apple_row[1], cfg.apple_col[1]) == (9, 0, 5, 10)
    assert cfg.perturb_prob == 0.25 and cfg.fear_radius == 5 and cfg.max_steps == 150


def test_header_is_plain_c_and_ctypes_mirrors_match(tmp_path):
    """include/gridworld_b200.h compiles as C99 (the boundary a cgo / JNI / ctypes binding would include), and every
    struct mirrored in _native.py has the size and field offsets the C compiler gives it."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("gcc not found")
    mirrors = {"gw_config": N.GwConfig, "gw_io": N.GwIO, "gw_stats": N.GwStats, "gw_actor_weights": N.GwActorWeights,
               "gw_replay_view": N.GwReplayView, "gww_config": N.GwwConfig, "gww_env_state": N.GwwEnvState}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "gridworld_b200.h")}"',
             'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in mirrors.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.BatchedGridWorld("Level 3", num_envs=4)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.CustomMAEnv()
    lib = N.load()
    cfg = N.build_config(pkg.builtin_scenario("Level 3"), num_envs=4)
    h = C.c_void_p()
    assert lib.gw_create(C.byref(cfg), C.byref(h)) == N.GW_ENODEV
    assert b"no CPU path" in lib.gw_last_error(None)
    bad = N.build_config(pkg.builtin_scenario("Level 3"), num_envs=4, n_agents=9)
    assert lib.gw_create(C.byref(bad), C.byref(h)) == N.GW_EINVAL


def test_product_does_not_import_oracle():
    pkg_dir = os.path.join(ROOT, "marl_responsible_nav_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "gridworld_oracle" not in src and "c_oracle" not in src and "gw_oracle" not in src, f


def test_builtin_scenarios_match_reference_tables(golden_dir):
    t = np.load(os.path.join(golden_dir, "scenario_tables.npz"))
    for name, tag in (("Level 3", "Level3"), ("Level 5", "Level5")):
        sc = pkg.builtin_scenario(name)
        assert np.array_equal(sc.region, t[tag + "_region"]) and sc.n_agents == int(t[tag + "_n_agents"])
        assert np.array_equal(np.array(sc.policy_keys)[sc.policy_index], t[tag + "_policy_map"])
        assert np.array_equal(sc.mdr_action, t[tag + "_mdr_action"])
        for i, k in enumerate(int(k) for k in t[tag + "_policy_keys"]):
            sw, dw = sc.policies[sc.policy_keys.index(k)]
            assert np.array_equal(pkg.policy_probs(sw, dw), t[tag + "_policy_base"][i])
            assert np.array_equal(pkg.policy_probs(sw, None), t[tag + "_policy_perturbed"][i])
    gm = pkg.builtin_scenario("GameMap")
    assert np.array_equal(gm.region, t["GameMap_region"]) and gm.n_agents == 3
    assert len(pkg.builtin_scenario("Level 3").active_cells()) == 72


def test_scenario_json_loader_roundtrip(tmp_path):
    """load_scenario_json reads the reference's JSON format: write Level 3 in that format and read it back."""
    import json
    sc = pkg.builtin_scenario("Level 3")

    def box(mask):
        rows, cols = np.where(mask)
        return {"slicex": [int(rows.min()), int(rows.max()) + 1, 0], "slicey": [int(cols.min()), int(cols.max()) + 1, 0]}

    policies = {}
    for i, k in enumerate(sc.policy_keys):          # bounding boxes in key order: later keys overwrite earlier ones
        policies[f"{k:02d}"] = dict(box(sc.policy_index == i) if i else box(sc.region >= 0),
                                    stepWeights=sc.policies[i][0], directionWeights=sc.policies[i][1])
    mdrs = {"00": dict(box(sc.region >= 0), mdr=0)}
    for n, act in enumerate((4, 3, 1, 2), start=1):
        mdrs[f"{n:02d}"] = dict(box(sc.mdr_action == act), mdr=act)
    doc = {"Level 3": {"AgentLocations": [], "N_Agents": 4, "Policies": policies, "MdRs": mdrs,
                       "Map": {"OneWays": [], "Walls": [], "Region": sc.region.astype(float).tolist()}}}
    p = tmp_path / "Scenarios.json"
    p.write_text(json.dumps(doc))
    back = pkg.load_scenario_json(str(p), "Level 3")
    assert np.array_equal(back.region, sc.region) and back.policies == sc.policies
    assert np.array_equal(back.policy_index, sc.policy_index) and back.n_agents == 4
    # the four corner cells belong to two MdR strips; JSON order decides, as in the reference
    assert (back.mdr_action != sc.mdr_action).sum() <= 4
    doc["Level 3"]["Map"]["Walls"] = [[[0, 0], [0, 1]]]
    p.write_text(json.dumps(doc))
    with pytest.raises(ValueError):
        pkg.load_scenario_json(str(p), "Level 3")


def test_reference_json_loader_live():
    ref = "/root/reference/custom/Scenarios.json"
    if not os.path.exists(ref):
        pytest.skip("reference not mounted")
    for name in ("Level 3", "Level 5", "GameMap"):
        a, b = pkg.load_scenario_json(ref, name), pkg.builtin_scenario(name)
        assert np.array_equal(a.region, b.region) and np.array_equal(a.mdr_action, b.mdr_action)
        assert a.n_agents == b.n_agents
        assert np.array_equal(np.array(a.policy_keys)[a.policy_index], np.array(b.policy_keys)[b.policy_index])
        for k in a.policy_keys:
            assert a.policies[a.policy_keys.index(k)] == b.policies[b.policy_keys.index(k)]
