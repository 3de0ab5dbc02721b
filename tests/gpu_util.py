import numpy as np


def first_diff(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.size != b.size:
        return "size %d != %d" % (a.size, b.size)
    d = np.flatnonzero(a != b)
    if d.size == 0:
        return "equal"
    i = int(d[0])
    return "%d mismatches of %d, first at %d: got %s want %s" % (d.size, a.size, i, a[i:i + 8].tolist(), b[i:i + 8].tolist())


def assert_same(a, b, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.size == b.size and np.array_equal(a, b), what + " " + first_diff(a, b)
