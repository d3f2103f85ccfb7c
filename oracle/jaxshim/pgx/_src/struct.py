import dataclasses


def dataclass(cls):
    cls = dataclasses.dataclass(cls)
    cls.replace = lambda self, **kw: dataclasses.replace(self, **kw)
    return cls
