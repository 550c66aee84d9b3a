from .metrics import MetricsSink, get_sink, set_sink

__all__ = ["MetricsSink", "get_sink", "set_sink"]
