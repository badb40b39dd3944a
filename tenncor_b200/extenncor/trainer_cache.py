"""Checkpoint / resume of a training session (reference: extenncor/trainer_cache.py:1-131).

A checkpoint is a pair of numbered files in `<cacheroot>/<name>/`: `session_<hex id>.onnx` (the graphs, in the reference's
ONNX dialect) and `env_<hex id>.bkup` (whatever the environment itself needs: counters, replay buffer). The highest id wins on
start-up; `clean=True` ignores what is there. Where the reference finds its handles again by querying the restored context
(`tc.Statement.find`, the json query engine that SURVEY.md §2 leaves out of scope), the session file here carries them as named
ids: `handles()` names the tensors an environment needs back, `restore(handles)` receives them."""
import abc
import os

import tenncor_b200 as tc

_default_cachedir = "/tmp"
_sess_prefix, _sess_ext = "session_", ".onnx"
_env_prefix, _env_ext = "env_", ".bkup"


def _id_cachefile(fpath, prefix, ext):
    """the hex id in `<prefix><id><ext>`, None for any other file (trainer_cache.py:14-23)"""
    fname = os.path.basename(fpath)
    if os.path.isfile(fpath) and fname.startswith(prefix) and fname.endswith(ext):
        try:
            return int(fname[len(prefix):len(fname) - len(ext)], 16)
        except ValueError:
            return None
    return None


def _latest(dirpath, prefix, ext):
    ids = [_id_cachefile(os.path.join(dirpath, el), prefix, ext) for el in os.listdir(dirpath)]
    return max([i for i in ids if i is not None], default=0)


class SessionCache:
    """CtxCache (trainer_cache.py:25-65): numbered model files"""

    def __init__(self, cache_dir=_default_cachedir):
        self.cache_dir = cache_dir
        self.cur_id = _latest(cache_dir, _sess_prefix, _sess_ext)

    def path(self, file_id):
        return os.path.join(self.cache_dir, _sess_prefix + hex(file_id)[2:] + _sess_ext)

    def backup(self, roots, handles):
        fpath = self.path(self.cur_id + 1)
        if tc.save_to_file(fpath, list(roots), dict(handles)):
            self.cur_id += 1
            return True
        return False

    def recover(self):
        """{handle name: tensor} of the newest session file, None when there is none"""
        fpath = self.path(self.cur_id)
        if self.cur_id == 0 or not os.path.isfile(fpath):
            return None
        _, ids = tc.load_model_ids(fpath)
        return ids


class EnvManager(metaclass=abc.ABCMeta):
    """trainer_cache.py:67-131. Subclasses build their graphs in `default_init`, name what they need back in `handles()` /
    `roots()`, and save / restore their own state in `_backup_env` / `_recover_env`."""

    def __init__(self, name, default_init=None, clean=False, cacheroot=_default_cachedir):
        self.dirpath = os.path.join(cacheroot, name)
        os.makedirs(self.dirpath, exist_ok=True)
        self.session_cache = SessionCache(self.dirpath)
        self.env_id = _latest(self.dirpath, _env_prefix, _env_ext)
        self.recovered = False
        if not clean:
            try:
                handles = self.session_cache.recover()
                if handles is not None:
                    self.restore(handles)
                    if self._recover_env(self.env_path(self.env_id)):
                        self.recovered = True
                        return
            except Exception as e:  # a damaged checkpoint must not stop a fresh start (trainer_cache.py:110-111)
                print("recovery error: {}".format(e))
        if default_init is not None:
            default_init()

    def env_path(self, file_id):
        return os.path.join(self.dirpath, _env_prefix + hex(file_id)[2:] + _env_ext)

    def backup(self):
        if self.session_cache.backup(self.roots(), self.handles()) and self._backup_env(self.env_path(self.env_id + 1)):
            self.env_id += 1
            return True
        return False

    @abc.abstractmethod
    def roots(self):
        """the graphs a checkpoint must hold"""

    @abc.abstractmethod
    def handles(self):
        """{name: tensor} of everything `restore` needs back"""

    @abc.abstractmethod
    def restore(self, handles):
        """take the tensors of a loaded session"""

    @abc.abstractmethod
    def _backup_env(self, fpath):
        """Backup environment settings"""

    @abc.abstractmethod
    def _recover_env(self, fpath):
        """Recover environment settings"""
