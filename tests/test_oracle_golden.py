"""Pins the CPU oracle (oracle/ftl_oracle.c) against traces of the unmodified reference.

The traces in tests/golden/ were exported by oracle/gen_golden.py from the reference's own
Game.reset/Game.step.  The oracle must reproduce every integer AND every float bit for bit
(positions, headings, speeds, rewards, tracker history, corridor); ray distances to 1e-6 relative
(the reference evaluates them through BLAS calls whose rounding is machine specific).
"""
import numpy as np
import pytest

import parity
from oracle_py import OracleEnv

FILES = parity.golden_files()


@pytest.mark.parametrize("path", FILES, ids=[p.split("/")[-1][:-4] for p in FILES])
def test_oracle_reproduces_reference_trace(path):
    d, meta = parity.load_trace(path)
    gc = parity.config_for(meta, _route_len=len(d["scen_route"]), _n_static=len(d["scen_static_rects"]))
    env = OracleEnv(gc, 1)
    env.upload_scenarios(parity.pool_for(d, gc))
    T, outliers = parity.replay(env, d, gc, float_rtol=0.0, ray_rtol=1e-6)
    assert T == meta["n_env_steps"]
    parity.check_final_arrays(env.get_state(), d, gc)
    assert int(env.get_state().env[0]["overflow"]) == 0


def test_golden_fixtures_present():
    assert len(FILES) >= 5
