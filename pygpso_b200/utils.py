"""
Small shared helpers: point labels, file-extension constants, JSON / directory helpers, logger setup.

Mirrors the public names of the reference's ``gpso/utils.py`` (``PointLabels`` :17-25, ``make_dirs`` :28-38,
``load_json`` :41-53, ``set_logger`` :56-84, extension constants :10-14) so user code keeps importing the same things.
"""
import json
import logging
import os
from enum import Enum, unique

LOG_DATETIME_FORMAT = "%Y-%m-%d %H:%M:%S"
LOG_EXT = ".log"
JSON_EXT = ".json"
PKL_EXT = ".pkl"
H5_EXT = ".h5"


@unique
class PointLabels(Enum):
    """State of a leaf / point: no score yet, objective evaluated at the centre, or score predicted by the GP."""

    not_assigned = 0
    evaluated = 1
    gp_based = 2


def make_dirs(path):
    """Create ``path`` (and parents); an existing directory only logs a warning, like the reference."""
    try:
        os.makedirs(path)
    except OSError as error:
        logging.warning(f"{path} could not be created: {error}")


def load_json(filename):
    with open(filename, "r") as handle:
        return json.load(handle)


def set_logger(log_filename=None, log_level=logging.INFO):
    """Console (and optionally file) logging on the root logger with the reference's line format."""
    formatter = logging.Formatter("[%(asctime)s] %(levelname)s: %(message)s", LOG_DATETIME_FORMAT)
    root = logging.getLogger()
    root.setLevel(log_level)
    root.handlers = []
    handlers = [logging.StreamHandler()]
    if log_filename is not None:
        if not log_filename.endswith(LOG_EXT):
            log_filename += LOG_EXT
        handlers.append(logging.FileHandler(log_filename))
    for handler in handlers:
        handler.setFormatter(formatter)
        handler.setLevel(log_level)
        root.addHandler(handler)
