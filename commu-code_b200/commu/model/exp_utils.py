"""Run logging: a per-rank log file under the work dir plus optional console echo (the role of
the reference's commu/model/exp_utils.py:logging_config)."""
import logging
import os


def logging_config(folder=None, name=None, level=logging.DEBUG, console_level=logging.INFO, console=True):
    name = name or "train"
    folder = folder or os.getcwd()
    os.makedirs(folder, exist_ok=True)
    root = logging.getLogger()
    for h in list(root.handlers):
        root.removeHandler(h)
    path = os.path.join(folder, name + ".log")
    print("All Logs will be saved to {}".format(path))
    root.setLevel(level)
    fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    fh = logging.FileHandler(path)
    fh.setLevel(level)
    fh.setFormatter(fmt)
    root.addHandler(fh)
    if console:
        ch = logging.StreamHandler()
        ch.setLevel(console_level)
        ch.setFormatter(fmt)
        root.addHandler(ch)
    return folder
