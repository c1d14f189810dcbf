"""Stand-in for the few things the reference's scripts use of path==15.0.0 (`from path import Path`, reference
README.md:71; run-testing.py:9,33,61-65,76-77): a str subclass with '/', files(pattern), dirs(), listdir(), makedirs_p().
Part of the environment overlay in compat/ (see sitecustomize.py), not of the dvmvs package."""
import fnmatch
import os


class Path(str):
    def __truediv__(self, other):
        return Path(os.path.join(self, str(other)))

    def __rtruediv__(self, other):
        return Path(os.path.join(str(other), self))

    def __add__(self, other):
        return Path(str.__add__(self, str(other)))

    def _entries(self, want_dir, pattern):
        names = sorted(os.listdir(self))
        keep = [n for n in names if os.path.isdir(os.path.join(self, n)) == want_dir]
        if pattern is not None:
            keep = [n for n in keep if fnmatch.fnmatch(n, pattern)]
        return [Path(os.path.join(self, n)) for n in keep]

    def files(self, pattern=None):
        return self._entries(False, pattern)

    def dirs(self, pattern=None):
        return self._entries(True, pattern)

    def listdir(self, pattern=None):
        return [Path(os.path.join(self, n)) for n in sorted(os.listdir(self)) if pattern is None or fnmatch.fnmatch(n, pattern)]

    def exists(self):
        return os.path.exists(self)

    def isdir(self):
        return os.path.isdir(self)

    def makedirs_p(self):
        os.makedirs(self, exist_ok=True)
        return self

    @property
    def name(self):
        return Path(os.path.basename(self))

    @property
    def parent(self):
        return Path(os.path.dirname(self))
