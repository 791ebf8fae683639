"""Minimal global flag registry with the interface of the reference's vendored `tf_flags`
(sqair/tf_flags.py:50-155): `DEFINE_{string,integer,boolean,bool,float}` register argparse
arguments on one global parser; `FLAGS.<name>` parses known args lazily on first access and can be
assigned to.  Re-implemented (no TensorFlow import)."""
import argparse as _argparse

_global_parser = _argparse.ArgumentParser(allow_abbrev=False)
_defined = set()


class _FlagValues(object):
    def __init__(self):
        self.__dict__['__flags'] = {}
        self.__dict__['__parsed'] = False

    def _parse_flags(self, args=None):
        result, unparsed = _global_parser.parse_known_args(args=args)
        for name, val in vars(result).items():
            self.__dict__['__flags'][name] = val
        self.__dict__['__parsed'] = True
        return unparsed

    def __getattr__(self, name):
        if not self.__dict__['__parsed']:
            self._parse_flags(args=[])
        flags = self.__dict__['__flags']
        if name not in flags:
            # a flag defined after the first parse: pick up its default
            if name in _defined:
                self._parse_flags(args=[])
                flags = self.__dict__['__flags']
            if name not in flags:
                raise AttributeError(name)
        return flags[name]

    def __setattr__(self, name, value):
        if not self.__dict__['__parsed']:
            self._parse_flags(args=[])
        self.__dict__['__flags'][name] = value


FLAGS = _FlagValues()


def _define(name, default, doc, typ):
    if name in _defined:            # configs may be imported twice (by path and by module name)
        return
    _defined.add(name)
    _global_parser.add_argument('--' + name, default=default, help=doc, type=typ)


def DEFINE_string(name, default, doc):
    _define(name, default, doc, str)


def DEFINE_integer(name, default, doc):
    _define(name, default, doc, int)


def DEFINE_float(name, default, doc):
    _define(name, default, doc, float)


def DEFINE_boolean(name, default, doc):
    if name in _defined:
        return
    _defined.add(name)
    _global_parser.add_argument('--' + name, nargs='?', const=True, default=default, help=doc,
                                type=lambda v: v.lower() in ('true', 't', '1'))
    _global_parser.add_argument('--no' + name, action='store_false', dest=name.replace('-', '_'))


DEFINE_bool = DEFINE_boolean
