"""argparse subclass understanding the subset of configargparse the reference uses
(config_parser.py:5-10): `is_config_file=True` options whose files hold `key = value`
lines, bare `key` lines for store_true flags, and `#` comments."""
import argparse


class ArgumentParser(argparse.ArgumentParser):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._config_dests = []

    def add_argument(self, *names, **kw):
        if kw.pop("is_config_file", False):
            act = super().add_argument(*names, **kw)
            self._config_dests.append(act.dest)
            return act
        return super().add_argument(*names, **kw)

    @staticmethod
    def _file_to_argv(path):
        out = []
        with open(path) as f:
            for line in f:
                line = line.split("#", 1)[0].strip() if line.strip().startswith("#") else line.strip()
                if not line:
                    continue
                if "=" in line:
                    k, v = line.split("=", 1)
                    out += ["--" + k.strip(), v.strip()]
                else:
                    out.append("--" + line)
        return out

    def parse_known_args(self, args=None, namespace=None):
        import sys
        args = list(sys.argv[1:] if args is None else args)
        pre, _ = super().parse_known_args(args, None)
        file_argv = []
        for dest in self._config_dests:
            path = getattr(pre, dest, None)
            if path:
                file_argv += self._file_to_argv(path)
        # command line overrides files; later files override earlier ones
        return super().parse_known_args(file_argv + args, namespace)
