"""Logging helpers with the reference's names and file formats (reference ``utils.py``):
``MetricLogger`` writes the ``valid{id}.csv`` / ``test{id}.csv`` files of kgat.py:86-91,
``creat_log_id`` picks the first unused ``log{n}.log`` and ``logging_config`` routes the root
logger to ``<folder>/<name>.log`` (+ console)."""
from __future__ import annotations

import csv
import logging
import os


class MetricLogger(object):
    def __init__(self, attr_names, parse_formats, save_path):
        self._names = list(attr_names)
        self._formats = dict(zip(attr_names, parse_formats))
        self._file = open(save_path, 'w')
        self._csv = csv.writer(self._file)
        self._csv.writerow(self._names)
        self._file.flush()

    def log(self, **kwargs):
        self._csv.writerow([self._formats[n] % kwargs[n] for n in self._names])
        self._file.flush()

    def close(self):
        self._file.close()


def creat_log_id(dir_path):
    n = 0
    while os.path.exists(os.path.join(dir_path, 'log{:d}.log'.format(n))):
        n += 1
    return n


def logging_config(folder=None, name=None, level=logging.DEBUG, console_level=logging.DEBUG, no_console=True):
    name = name or "kgat"
    folder = folder or os.path.join(os.getcwd(), name)
    os.makedirs(folder, exist_ok=True)
    root = logging.getLogger()
    for handler in list(root.handlers):
        root.removeHandler(handler)
    logpath = os.path.join(folder, name + ".log")
    print("All Logs will be saved to %s" % logpath)
    root.setLevel(level)
    fmt = logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s')
    fh = logging.FileHandler(logpath)
    fh.setLevel(level)
    fh.setFormatter(fmt)
    root.addHandler(fh)
    if not no_console:
        ch = logging.StreamHandler()
        ch.setLevel(console_level)
        ch.setFormatter(fmt)
        root.addHandler(ch)
    return folder
