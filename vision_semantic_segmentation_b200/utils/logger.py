"""Run logger with the shape of the reference's ``MyLogger`` (``src/utils/logger.py:11-91``):
messages go to stdout and to ``<save_dir>/version_<n>/log.txt``; ``.save_dir`` is the versioned folder."""
import logging
import os
import sys


class MyLogger(object):
    def __init__(self, name, save_dir="", version=None, use_timestamp=True):
        logger = logging.getLogger("%s.%d" % (name, id(self)))
        logger.setLevel(logging.DEBUG)
        logger.propagate = False
        fmt = logging.Formatter("%(asctime)s " + name + " %(levelname)s: %(message)s")
        console = logging.StreamHandler(stream=sys.stdout)
        console.setFormatter(fmt)
        logger.addHandler(console)
        if save_dir:
            if version is None:
                version = self._next_version(save_dir)
            save_dir = os.path.join(save_dir, "version_" + str(version))
            os.makedirs(save_dir, exist_ok=True)
            stem = "log"
            if use_timestamp:
                import socket
                import time
                stem += "." + time.strftime("%m-%d_%H-%M-%S") + "." + socket.gethostname()
            fh = logging.FileHandler(os.path.join(save_dir, stem + ".txt"))
            fh.setFormatter(fmt)
            logger.addHandler(fh)
        self.logger = logger
        self.save_dir = save_dir

    def log(self, msg, level="info"):
        if level not in ("info", "debug", "warning", "critical"):
            raise NotImplementedError
        getattr(self.logger, level)(msg)

    @staticmethod
    def _next_version(save_dir):
        os.makedirs(save_dir, exist_ok=True)
        taken = [int(d.split("_")[1]) for d in os.listdir(save_dir)
                 if d.startswith("version_") and os.path.isdir(os.path.join(save_dir, d))
                 and d.split("_")[1].isdigit()]
        return max(taken) + 1 if taken else 0
