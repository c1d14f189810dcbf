"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for path==15.0.0 (reference README.md:71): a str
subclass with '/', '+', files(pattern), listdir(), dirs(), makedirs_p()."""
import fnmatch
import os


class Path(str):
    def __truediv__(self, other):
        return Path(os.path.join(self, other))

    def __add__(self, other):
        return Path(str.__add__(self, other))

    def files(self, pattern=None):
        out = [Path(os.path.join(self, f)) for f in os.listdir(self) if os.path.isfile(os.path.join(self, f))]
        if pattern is not None:
            out = [f for f in out if fnmatch.fnmatch(os.path.basename(f), pattern)]
        return out

    def dirs(self, pattern=None):
        out = [Path(os.path.join(self, f)) for f in os.listdir(self) if os.path.isdir(os.path.join(self, f))]
        if pattern is not None:
            out = [f for f in out if fnmatch.fnmatch(os.path.basename(f), pattern)]
        return out

    def listdir(self):
        return [Path(os.path.join(self, f)) for f in os.listdir(self)]

    def makedirs_p(self):
        os.makedirs(self, exist_ok=True)
        return self

    def exists(self):
        return os.path.exists(self)
