"""CPU: the R glue (rsplash_b200/rglue/rglue.cpp) compiles and links against the C ABI, registers the routines the
R side .Call()s, and turns a failed context creation into an R error instead of a crash (no GPU in this container)."""
import numpy as np
import pytest

from tests import rglue_harness as rh


def test_glue_builds_and_registers_its_routines():
    L = rh.lib()
    table = {L.stub_routine_name(i).decode(): L.stub_routine_nargs(i) for i in range(L.stub_n_routines())}
    assert table == {"splash_grid_run_R": 15, "splash_point_run_R": 15, "splash_grid_submit_R": 16, "splash_grid_wait_R": 1,
                     "splash_unswc_grid_R": 4, "splash_month2day_linear_R": 4, "splash_release_R": 0}


def test_bad_shapes_raise_r_errors_before_touching_the_device():
    z = np.zeros
    args = [rh.r_matrix(z((5, 3))), rh.r_matrix(z((5, 3))), rh.r_matrix(z((4, 3))),             # pn has the wrong shape
            *(rh.r_matrix(z(3)) for _ in range(4)), rh.r_matrix(z((6, 3))), rh.r_matrix(z((1, 3))), rh.r_matrix(z(3)),
            rh.r_int(z(5)), rh.r_int(z(5)), rh.r_int(z(5)), rh.r_lgl(True), rh.r_int([0])]
    with pytest.raises(rh.RError, match="`pn` has 12 elements, expected 15"):
        rh.dot_call("splash_grid_run_R", *args)
    assert rh.lib().stub_protect_depth() == 0


def _good_args(nc=3, nd=5):
    z = np.zeros
    return [rh.r_matrix(z((nd, nc))), rh.r_matrix(z((nd, nc))), rh.r_matrix(z((nd, nc))),
            *(rh.r_matrix(z(nc)) for _ in range(4)), rh.r_matrix(z((6, nc))), rh.r_matrix(z((1, nc))), rh.r_matrix(z(nc)),
            rh.r_int(z(nd)), rh.r_int(z(nd)), rh.r_int(z(nd)), rh.r_lgl(True), rh.r_int([0])]


def test_every_argument_is_length_and_type_checked():
    """a scalar resolution, an integer matrix or a short doy vector must be R errors, not wild reads."""
    for pos, bad, msg in ((9, rh.r_matrix(np.zeros(1)), "`resolution` has 1 elements, expected 3"),
                          (4, rh.r_matrix(np.zeros(2)), "`elev` has 2 elements"),
                          (11, rh.r_int(np.zeros(4)), "`doy` has 4 elements, expected 5"),
                          (7, rh.r_int(np.zeros(18)), "`soil` must be a double"),
                          (10, rh.r_matrix(np.zeros(5)), "`year` must be an integer vector"),
                          (8, rh.r_matrix(np.zeros((2, 3))), "`Au` must be"),
                          (14, rh.r_matrix(np.zeros(0)), "`device` must name")):
        args = _good_args()
        args[pos] = bad
        with pytest.raises(rh.RError, match=msg):
            rh.dot_call("splash_grid_run_R", *args)
        assert rh.lib().stub_protect_depth() == 0
    args = _good_args()
    args[7] = rh.r_matrix(np.zeros(5))  # soil_data of splash.point needs six values
    with pytest.raises(rh.RError, match="`sw_in` has 15 elements, expected 5|`soil_data`"):
        rh.dot_call("splash_point_run_R", *args)
    with pytest.raises(rh.RError, match="nothing was submitted"):
        rh.dot_call("splash_grid_wait_R", rh.r_matrix(np.array([-1.0])))


def test_without_a_gpu_the_routine_stops_with_the_library_message():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by tests/test_rglue_gpu.py")
    with pytest.raises(rh.RError, match="libsplash_cuda"):
        rh.dot_call("splash_month2day_linear_R", rh.r_matrix(np.zeros((12, 4))), rh.r_int(np.arange(12) * 30), rh.r_int([365]), rh.r_int([0]))
    assert rh.lib().stub_protect_depth() == 0
