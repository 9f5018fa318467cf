"""``VelocityVectorField`` -- event-based visual flow (Benosman et al. 2014) as the reference implements it
(stereovis/spiking/algorithms/vvf.py:5-110), with the per-event plane fits on the GPU
(``hsm_fit_velocity_field``).  Same constructor arguments, methods and return shapes.  The reference's only user,
``OpticalFlowPixelCorrection.adjust``, computes this field and then does not use it
(optical_flow_correction.py:69, FIXME); it is provided for callers that want the field itself."""
import ctypes

import numpy as np

from . import _capi


class VelocityVectorField(object):
    def __init__(self, time_interval=20, neighbourhood_size=(3, 3), rejection_threshold=0.005,
                 convergence_threshold=1e-5, max_iter_steps=5, min_num_events_in_timespace_interval=10, device=0):
        self.delta_t = time_interval
        self.neighbourhood_size = neighbourhood_size
        self.rejection_threshold = rejection_threshold
        self.convergence_threshold = convergence_threshold
        self.max_iter_steps = max_iter_steps
        self.min_num_events_in_timespace_interval = min_num_events_in_timespace_interval
        self.device = int(device)

    def _fit_velocity_field(self, events):
        """``events``: (N, 3) array of (t, x, y), sorted by time -> (N, 2) x, y velocity components (vvf.py:32-82)."""
        events = np.asarray(events, dtype=np.float64)
        n = events.shape[0]
        velocities = np.zeros((n, 2))
        if n == 0:
            return velocities
        t, x, y = (np.ascontiguousarray(events[:, k]) for k in range(3))
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _capi.check(_capi.lib().hsm_fit_velocity_field(
            n, ptr(t), ptr(x), ptr(y), float(self.delta_t), float(self.neighbourhood_size[0]),
            float(self.neighbourhood_size[1]), float(self.rejection_threshold), float(self.convergence_threshold),
            int(self.max_iter_steps), int(self.min_num_events_in_timespace_interval), ptr(velocities), self.device))
        return velocities

    def fit_velocity_field(self, events, assume_sorted=False, concatenate_polarity_groups=True):
        """``events``: (N, 4) array of (t, x, y, polarity); polarity 0 = "positive" group (vvf.py:84-110)."""
        events = np.asarray(events)
        if not assume_sorted:
            events = events[np.argsort(events[:, 0], kind="stable")]
        positive = events[:, 3] == 0
        positive_velocities = self._fit_velocity_field(events[positive][:, :-1])
        negative_velocities = self._fit_velocity_field(events[~positive][:, :-1])
        if concatenate_polarity_groups:
            return np.vstack([positive_velocities, negative_velocities])
        return {"positive": positive_velocities, "negative": negative_velocities}
