"""Constants-only stand-in for the author's un-vendored `spatial_maths` package (TEST INFRASTRUCTURE).

The reference imports `spatial_maths.camera_model_parameters as pidx`
(camera_model/distorted_camera_model.py:3) but the package is neither in pyproject.toml nor in
poetry.lock.  Only 16 index constants and `make_camera_parameters` are used; their layout is pinned
by tests/camera_model/test_distorted_camera_model.py:13-30,127,163,204-206 and the Jacobian column
order camera_model/distorted_camera_model.py:364-383.
"""
