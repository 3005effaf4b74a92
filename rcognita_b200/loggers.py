"""Console / CSV loggers with the reference's row formats (rcognita/loggers.py:36-94): one class per system,
``print_sim_step`` (grid table) and ``log_data_row`` (CSV append).  Host-side IO for ONE environment's row per call;
for a batch, pass the values of the environment you want to follow (the presets log environment 0), or use
``log_data_rows`` to append the same columns plus a leading ``env`` column for several environments at once.

Column orders (they define the trajectory file format a reference user's tooling expects):
  3wrobotNI : t, x, y, alpha, stage_obj, accum_obj, v, omega
  3wrobot   : t, x, y, alpha, v, omega, stage_obj, accum_obj, F, M
  2tank     : t, h1, h2, p, stage_obj, accum_obj
"""
from __future__ import annotations

import csv

try:
    from tabulate import tabulate
except ImportError:                                       # pragma: no cover
    tabulate = None


def _f(v):
    """Scalars of any flavour (python, numpy, 0-d / 1-element tensors) -> float."""
    try:
        return float(v)
    except (TypeError, ValueError):
        return float(v.reshape(-1)[0])


class Logger:
    """Interface (rcognita/loggers.py:12-34)."""
    HEADER: list = []
    FORMAT: tuple = ()

    def _row(self, *args):
        raise NotImplementedError

    def print_sim_step(self, *args):
        row = self._row(*args)
        if tabulate is not None:
            print(tabulate([self.HEADER, row], floatfmt=self.FORMAT, headers='firstrow', tablefmt='grid'))
        else:
            print('  '.join(f'{h}={format(v, f)}' for h, v, f in zip(self.HEADER, row, self.FORMAT)))

    def log_data_row(self, datafile, *args):
        with open(datafile, 'a', newline='') as outfile:
            csv.writer(outfile).writerow(self._row(*args))

    def log_data_rows(self, datafile, env_ids, rows):
        """Batched extension: ``rows[i]`` are the ``log_data_row`` arguments of environment ``env_ids[i]``."""
        with open(datafile, 'a', newline='') as outfile:
            w = csv.writer(outfile)
            for e, args in zip(env_ids, rows):
                w.writerow([int(e)] + self._row(*args))


    def log_trajectory(self, datafile, rows):
        """Append a whole logged trajectory (``ClosedLoopEngine.trajectory(e)``: rows already in this logger's
        column order) to a CSV in the reference's row format."""
        with open(datafile, 'a', newline='') as outfile:
            w = csv.writer(outfile)
            for r in rows:
                w.writerow([float(v) for v in r])


class Logger3WRobot(Logger):
    """rcognita/loggers.py:36-56."""
    HEADER = ['t [s]', 'x [m]', 'y [m]', 'alpha [rad]', 'v [m/s]', 'omega [rad/s]', 'stage_obj', 'accum_obj', 'F [N]', 'M [N m]']
    FORMAT = ('8.3f', '8.3f', '8.3f', '8.3f', '8.3f', '8.3f', '8.1f', '8.1f', '8.3f', '8.3f')

    def _row(self, t, xCoord, yCoord, alpha, v, omega, stage_obj, accum_obj, action):
        return [_f(t), _f(xCoord), _f(yCoord), _f(alpha), _f(v), _f(omega), _f(stage_obj), _f(accum_obj),
                _f(action[0]), _f(action[1])]


class Logger3WRobotNI(Logger):
    """rcognita/loggers.py:58-76."""
    HEADER = ['t [s]', 'x [m]', 'y [m]', 'alpha [rad]', 'stage_obj', 'accum_obj', 'v [m/s]', 'omega [rad/s]']
    FORMAT = ('8.3f', '8.3f', '8.3f', '8.3f', '8.1f', '8.1f', '8.3f', '8.3f')

    def _row(self, t, xCoord, yCoord, alpha, stage_obj, accum_obj, action):
        return [_f(t), _f(xCoord), _f(yCoord), _f(alpha), _f(stage_obj), _f(accum_obj), _f(action[0]), _f(action[1])]


class Logger2Tank(Logger):
    """rcognita/loggers.py:78-94."""
    HEADER = ['t [s]', 'h1', 'h2', 'p', 'stage_obj', 'accum_obj']
    FORMAT = ('8.1f', '8.4f', '8.4f', '8.4f', '8.4f', '8.2f')

    def _row(self, t, h1, h2, p, stage_obj, accum_obj):
        return [_f(t), _f(h1), _f(h2), _f(p), _f(stage_obj), _f(accum_obj)]
