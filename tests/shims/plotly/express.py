"""plotly.express stand-in: ``histogram`` returns the arguments it was given (train_muse.py:1371-1378 logs the figure)."""


class _Qualitative:
    Plotly = ["#636EFA", "#EF553B", "#00CC96"]


class colors:  # noqa: N801
    qualitative = _Qualitative


def histogram(data_frame=None, **kwargs):
    return dict(kind="histogram", rows=0 if data_frame is None else len(data_frame), **kwargs)
