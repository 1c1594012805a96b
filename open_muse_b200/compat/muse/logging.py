"""``muse.logging`` for the drop-in package: the verbosity switches the training scripts call
(training/train_maskgit_imagenet.py:175-178; reference muse/logging.py:107-181) on a library root logger."""
import logging
import sys

_root = logging.getLogger("muse")
if not _root.handlers:
    _h = logging.StreamHandler(sys.stderr)
    _root.addHandler(_h)
    _root.setLevel(logging.WARNING)
    _root.propagate = False

DEBUG, INFO, WARNING, ERROR, CRITICAL = logging.DEBUG, logging.INFO, logging.WARNING, logging.ERROR, logging.CRITICAL


def get_logger(name=None):
    return logging.getLogger(name or "muse")


def get_verbosity():
    return _root.getEffectiveLevel()


def set_verbosity(verbosity):
    _root.setLevel(verbosity)


def set_verbosity_info():
    set_verbosity(INFO)


def set_verbosity_warning():
    set_verbosity(WARNING)


def set_verbosity_debug():
    set_verbosity(DEBUG)


def set_verbosity_error():
    set_verbosity(ERROR)
