"""Minimal stand-in for the `parameterized` package (absent from this image, no network): the one entry point
the reference's unittest uses, `@parameterized.expand(list_of_(name, arg)_tuples, skip_on_empty=True)`
(/root/reference/py/flash_helpers/test/test.py:73-99).  Only put on sys.path when the real package is missing."""
import re
import sys


class parameterized:  # noqa: N801  (name of the package's class)
    @staticmethod
    def expand(cases, skip_on_empty=False):
        cases = list(cases)

        def deco(fn):
            frame_locals = sys._getframe(1).f_locals  # the class body under construction
            for i, case in enumerate(cases):
                args = case if isinstance(case, (tuple, list)) else (case,)
                suffix = re.sub(r"\W+", "_", str(args[0])).strip("_")

                def test(self, _args=tuple(args)):
                    return fn(self, *_args)

                test.__name__ = f"{fn.__name__}_{i}_{suffix}"
                frame_locals[test.__name__] = test
            return None  # the template itself is not a test

        return deco
