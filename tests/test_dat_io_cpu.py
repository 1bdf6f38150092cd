"""On-disk formats (host code, no GPU): the reference's .dat files round-trip through the library's
reader / writer (RRT_Planner::load_DS_database rrt_planner.cpp:121-191; writers
stable_config_generator.cpp:463-468, rrt_planner.cpp:1444-1452)."""
import numpy as np
import pytest

import nao_wholebody_planning_b200 as nwb
import oracle_lib as OL

CD = OL.GOLDEN / "config_database"


@pytest.mark.parametrize("name", ["ssc_double_support.dat", "ssc_double_support_backup.dat"])
def test_reader_matches_the_shipped_files(name):
    rows = nwb.load_ds_database(CD / name)
    ref = OL.load_dat(CD / name)
    assert rows.shape == ref.shape and (rows == ref).all()


def test_solution_file_and_goal_files(tmp_path):
    sol = nwb.load_ds_database(OL.GOLDEN / "solutions" / "solution_path_second.dat")
    assert sol.shape == (106, 21)
    # goal files carry 22 columns (ergonomics index first, SCG:409): not 21-column configurations... the
    # reader takes the first 21 values of each line, exactly as load_DS_database's strtod loop would
    g = nwb.load_ds_database(CD / "ssc_double_support_goal_second.dat")
    assert g.shape == (6, 21)


def test_writer_reproduces_the_reference_format_byte_for_byte(tmp_path):
    src = CD / "ssc_double_support.dat"
    rows = nwb.load_ds_database(src)
    out = tmp_path / "db.dat"
    nwb.write_configs(out, rows)
    assert out.read_bytes() == src.read_bytes()          # "%g " per value, newline per row
    again = nwb.load_ds_database(out)
    assert (again == rows).all()


def test_missing_file_is_an_error_not_an_exit(tmp_path):
    with pytest.raises(nwb.NwbError):
        nwb.load_ds_database(tmp_path / "nope.dat")
    with pytest.raises(nwb.NwbError):
        nwb.write_configs(tmp_path / "no_such_dir" / "x.dat", np.zeros((1, 21)))
