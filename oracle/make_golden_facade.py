"""Generate ``tests/golden/f*.npz`` / ``r*.npz`` by running the REAL reference callers either side of the
hot path in this container: ``FramebasedStereoMatching`` (oracle/facade_ref.py) and
``generate_frames_from_spikes`` (oracle/frames_ref.py).  TEST INFRASTRUCTURE ONLY; run once from the repo
root (``python -m oracle.make_golden_facade``), outputs are committed."""
import sys

import numpy as np

from . import facade_cases as F
from . import facade_ref, flow_ref, frames_ref


def main():
    facade_class = facade_ref.load()
    for case in F.FACADE_CASES:
        lefts, rights, ts, info = F.facade_inputs(case)
        facade = facade_class((case["cols"], case["rows"]), case["levels"],
                              inputs={"left": lefts, "right": rights, "ts": ts})
        if case["priors"] == "online":         # the online call pattern: run_one_frame(left, right, prior, n_iter=k)
            for i in range(case["n"]):
                facade.run_one_frame(lefts[i], rights[i], info["priors"][i], n_iter=3 + i)
        else:
            facade.run(info)
        out = facade.get_output()
        np.savez_compressed(F.fixture_path(case), depth=out.astype(np.int8), timestamps=facade.get_timestamps(),
                            dtype=np.asarray(str(out.dtype)))
        print("%-36s %s %s" % (case["name"], out.shape, out.dtype), flush=True)
    _, rasterise = frames_ref.load()
    for case in F.RASTER_CASES:
        xs, ys, ts, zs = F.raster_events(case)
        arrays = {}
        for fill in (-1, "nan"):
            frames, stamps = rasterise(case["res"], xs, ys, ts, zs, start_time=case.get("start_time", 0),
                                       time_interval=case["interval"], pivots=case["pivots"], non_pixel_value=fill)
            arrays["frames_%s" % fill] = frames
            arrays["stamps_%s" % fill] = np.asarray(stamps)
        np.savez_compressed(F.fixture_path(case), **arrays)
        print("%-36s %s" % (case["name"], arrays["frames_-1"].shape), flush=True)
    field_class, correction_class = flow_ref.load()
    for case in F.FLOW_CASES:
        events = F.flow_events(case)
        field = field_class(time_interval=case["dt"], neighbourhood_size=(3, 3), rejection_threshold=0.005,
                            convergence_threshold=1e-5, max_iter_steps=5, min_num_events_in_timespace_interval=10)
        out = field.fit_velocity_field(events, assume_sorted=False, concatenate_polarity_groups=False)
        np.savez_compressed(F.fixture_path(case), positive=out["positive"], negative=out["negative"])
        fitted = int((out["positive"] != 0).any(axis=1).sum() + (out["negative"] != 0).any(axis=1).sum())
        print("%-36s %d events, %d fitted" % (case["name"], events.shape[0], fitted), flush=True)
    for case in F.ADJUST_CASES:
        events, frames = F.adjust_inputs(case)
        correction = correction_class(case["res"], events, buffer_pivots=case["pivots"],
                                      buffer_interval=case["interval"])
        adjusted = correction.adjust(frames, prior_nondata_value=-1, time_arrow=case["time_arrow"])
        np.savez_compressed(F.fixture_path(case), adjusted=adjusted, dtype=np.asarray(str(adjusted.dtype)))
        print("%-36s %s %s moved %d" % (case["name"], adjusted.shape, adjusted.dtype,
                                        int((adjusted != frames).sum())), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
