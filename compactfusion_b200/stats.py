"""API-compatible no-op stand-in for xfuser/compact/stats.py (diagnostics only; out of scope,
SURVEY.md section 2 row 10).  `log_stats=True` configs run, nothing is recorded."""


class StatsLogger:
    def log(self, *args, **kwargs):
        return None

    def clear(self):
        return None


_logger = StatsLogger()


def stats_log():
    return _logger


def stats_hello():
    return None


def stats_clear():
    _logger.clear()


def stats_verbose(*args, **kwargs):
    return None


def stats_verbose_steps(*args, **kwargs):
    return None


def plot_eigenvalues(*args, **kwargs):
    return None


def save_eigenvalues(*args, **kwargs):
    return None


def dump_err_vs_steps(*args, **kwargs):
    return None
