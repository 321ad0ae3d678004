"""Fixed-step headless driver: the calling convention of ``start_main_loop``
(src/viewer/mujoco_viewer.py:106-139) without the window, the renderer or the vsync pacing.

Per iteration, exactly as the reference: ``pos_new = step_function(model, data, dt=model.opt.timestep)``
(:113), ``simulation_time += model.opt.timestep`` (:114) and, when a logger is given and ``pos_new`` has three
components, ``logger.record(simulation_time, z, x, y)`` (:116-119).
"""
import numpy as np
import torch


class TrajectoryLog:
    """Device-side replacement for the reference's list-appending loggers
    (src/visualization/logger_base.py:22-32, data_logger.py:15-24): positions of the first ``n_sample``
    environments are written into a preallocated device buffer each step -- no host sync on the step path.
    After ``finish()`` the reference's attribute names are available for environment 0."""

    def __init__(self, steps, n_sample, device, dtype=torch.float64):
        self.buf = torch.empty((steps, n_sample, 3), dtype=dtype, device=device)
        self.t = np.zeros(steps)
        self.n = 0
        self.n_sample = n_sample
        self.times, self.x_positions, self.y_positions, self.z_positions = [], [], [], []

    def record(self, time_point, z_position, x_position=None, y_position=None):
        """Same argument order as DataLogger.record (z first).  Arguments may be floats (one environment)
        or device tensors of per-environment values."""
        i = self.n
        if i >= self.buf.shape[0]:
            raise IndexError("TrajectoryLog is full")
        for c, val in enumerate((x_position, y_position, z_position)):
            if val is None:
                self.buf[i, :, c] = 0
            elif torch.is_tensor(val):
                self.buf[i, :, c] = val.reshape(-1)[: self.n_sample]
            else:
                self.buf[i, :, c] = float(val)
        self.t[i] = time_point
        self.n += 1

    def window(self, count, time0, dt):
        """The next ``count`` rows as a contiguous ``[count, n_sample, 3]`` device view for a fused launch to fill
        (``trajectory=`` of the step functions); stamps their times ``time0 + (i+1)*dt`` as the per-frame loop would."""
        i = self.n
        if i + count > self.buf.shape[0]:
            raise IndexError("TrajectoryLog is full")
        self.t[i:i + count] = time0 + dt * np.arange(1, count + 1)
        self.n += count
        return self.buf[i:i + count]

    def finish(self):
        host = self.buf[: self.n].cpu().numpy()
        self.times = self.t[: self.n].tolist()
        self.x_positions, self.y_positions, self.z_positions = (host[:, 0, c].tolist() for c in range(3))
        return host

    def save_npz(self, path):
        """times [n] and positions [n, n_sample, 3] (x, y, z) of everything recorded so far."""
        np.savez(path, times=self.t[: self.n], positions=self.buf[: self.n].cpu().numpy())


def start_main_loop(model, data, step_function, steps, logger=None, substeps_per_launch=1):
    """Runs ``steps`` iterations of the reference's loop body; returns the accumulated simulation time.

    ``substeps_per_launch = K > 1`` fuses K iterations into each call of ``step_function`` (which must accept the
    ``substeps`` keyword); with a :class:`TrajectoryLog` the kernel itself records the sampled environments' position
    after every one of the K steps (``trajectory=``), so the log is the one the per-frame loop would have written."""
    simulation_time = 0.0
    K = int(substeps_per_launch)
    if K > 1:
        done = 0
        while done < int(steps):
            k = min(K, int(steps) - done)
            if isinstance(logger, TrajectoryLog):
                step_function(model, data, dt=model.opt.timestep, substeps=k,
                              trajectory=logger.window(k, simulation_time, model.opt.timestep))
            else:
                step_function(model, data, dt=model.opt.timestep, substeps=k)
            simulation_time += model.opt.timestep * k
            done += k
        return simulation_time
    for _ in range(int(steps)):
        pos_new = step_function(model, data, dt=model.opt.timestep)
        simulation_time += model.opt.timestep
        if logger is not None and pos_new is not None:
            if torch.is_tensor(pos_new) and pos_new.dim() == 2 and pos_new.shape[-1] == 3:
                logger.record(simulation_time, pos_new[:, 2], pos_new[:, 0], pos_new[:, 1])
            elif len(pos_new) == 3:
                x, y, z = pos_new
                logger.record(simulation_time, z, x, y)
    return simulation_time
