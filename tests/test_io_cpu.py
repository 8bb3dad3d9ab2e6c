"""Reference file formats (tests/slow_tests/test_experiments.py:178-191 pins
the reader on data/sample_macro.csv; here the same three rows are used)."""
import numpy as np

from mbrl_traffic_b200.io import (load_initial_conditions, read_macro_csv,
                                  write_macro_csv)

SAMPLE = "time,speed_0,speed_1,density_0,density_1\n0,1,2,3,4\n0.5,5,6,7,8\n1,9,10,11,12\n"


def test_read_reorders_to_density_first(tmp_path):
    p = tmp_path / "sample_macro.csv"
    p.write_text(SAMPLE)
    times, obs = read_macro_csv(str(p))
    np.testing.assert_almost_equal(times, [0, 0.5, 1])
    # reference import_data returns [speed | density]; the models need the swap
    np.testing.assert_almost_equal(obs, [[3, 4, 1, 2], [7, 8, 5, 6], [11, 12, 9, 10]])


def test_write_then_read_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    obs = rng.uniform(0, 1, (5, 12))
    times = np.arange(5) * 0.5
    p = tmp_path / "macro_0_0.csv"
    write_macro_csv(str(p), times, obs)
    header = p.read_text().splitlines()[0].split(",")
    assert header == ["time"] + ["speed_%d" % k for k in range(6)] + \
        ["density_%d" % k for k in range(6)]
    t2, obs2 = read_macro_csv(str(p))
    np.testing.assert_allclose(t2, times)
    np.testing.assert_allclose(obs2, obs, rtol=1e-15)


def test_initial_conditions_without_time_column(tmp_path):
    p = tmp_path / "60.csv"
    p.write_text("speed_0,speed_1,speed_10,density_0,density_1,density_10\n"
                 "30,20,10,0.0,0.1,0.2\n")
    obs = load_initial_conditions(str(p), min_density=1e-3)
    np.testing.assert_allclose(obs, [[1e-3, 0.1, 0.2, 30, 20, 10]])
