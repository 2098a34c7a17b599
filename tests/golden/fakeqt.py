"""numpy-backed stand-in for the slice of PyQt6 the reference's hot path touches.

PyQt6 is not installed in the build container, so ``import indexing_ui`` fails
there.  ``install()`` registers a fake ``PyQt6`` package so the UNMODIFIED
reference module imports, and gives QImage / QPixmap / QListWidget just enough
behaviour (documented Qt semantics) for
``MainWindow.run_segmentation_all``, ``_detect_defects_on_pix``,
``run_inspection``, ``update_grid_preview``, ``populate_thumbnails`` and
``import_grid`` to execute on numpy arrays.  Used ONLY by make_golden.py.

Semantics modelled:
  * Grayscale8 QImage <-> QPixmap round trips preserve bytes.
  * QImage.copy(x,y,w,h) pads with 0 outside the image; QPixmap.copy clips.
  * QImage.scaled(size) with an identical size returns the image unchanged;
    any other size raises (the hot path never resizes).
  * segmentation.qimage_to_gray_array on a mono image is the identity
     -- make_golden.py patches it to return the array.
"""
import sys
import types

import numpy as np


class _Meta(type):
    def __getattr__(cls, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return Stub

    def __int__(cls):
        return 256

    def __index__(cls):
        return 256

    def __or__(cls, o):
        return cls

    def __ror__(cls, o):
        return cls


class Stub(metaclass=_Meta):
    """Absorbs any UI call."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return Stub()

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return Stub()

    def __int__(self):
        return 256

    def __index__(self):
        return 256

    def __or__(self, o):
        return self

    def __ror__(self, o):
        return self

    def __bool__(self):
        return False

    def __iter__(self):
        return iter(())


class QSize:
    def __init__(self, w, h):
        self._w, self._h = int(w), int(h)

    def width(self):
        return self._w

    def height(self):
        return self._h

    def __eq__(self, o):
        return (self._w, self._h) == (o._w, o._h)


class QRect:
    def __init__(self, x=0, y=0, w=0, h=0):
        self._x, self._y, self._w, self._h = int(x), int(y), int(w), int(h)

    def x(self):
        return self._x

    def y(self):
        return self._y

    def width(self):
        return self._w

    def height(self):
        return self._h


class QImage:
    class Format:
        Format_Grayscale8 = 24
        Format_ARGB32 = 5

    def __init__(self, data=None, w=0, h=0, bpl=0, fmt=None):
        if isinstance(data, np.ndarray):
            self.arr = np.ascontiguousarray(data, dtype=np.uint8)
        elif data is None:
            self.arr = np.zeros((0, 0), np.uint8)
        else:
            assert fmt == QImage.Format.Format_Grayscale8 and bpl == w
            self.arr = np.frombuffer(bytes(data), np.uint8).reshape(int(h), int(w)).copy()

    def width(self):
        return self.arr.shape[1]

    def height(self):
        return self.arr.shape[0]

    def size(self):
        return QSize(self.arr.shape[1], self.arr.shape[0])

    def isNull(self):
        return self.arr.size == 0

    def copy(self, *a):
        if not a:
            return QImage(self.arr.copy())
        x, y, w, h = [int(v) for v in a]
        out = np.zeros((h, w), np.uint8)            # QImage.copy pads with 0
        H, W = self.arr.shape
        x0, y0, x1, y1 = max(0, x), max(0, y), min(W, x + w), min(H, y + h)
        if x1 > x0 and y1 > y0:
            out[y0 - y:y1 - y, x0 - x:x1 - x] = self.arr[y0:y1, x0:x1]
        return QImage(out)

    def scaled(self, *a, **k):
        size = a[0]
        if isinstance(size, QSize):
            tw, th = size.width(), size.height()
        else:
            tw, th = int(a[0]), int(a[1])
        if (tw, th) != (self.width(), self.height()):
            raise NotImplementedError('fake QImage.scaled only models the identical-size no-op')
        return self


class QPixmap:
    def __init__(self, arr=None):
        self.arr = None if arr is None else np.ascontiguousarray(arr, dtype=np.uint8)

    @staticmethod
    def fromImage(qimg):
        return QPixmap(qimg.arr.copy())

    def toImage(self):
        return QImage(self.arr.copy())

    def copy(self, x, y, w, h):
        H, W = self.arr.shape                        # QPixmap.copy clips
        x0, y0, x1, y1 = max(0, x), max(0, y), min(W, x + w), min(H, y + h)
        return QPixmap(self.arr[y0:y1, x0:x1].copy())

    def scaled(self, *a, **k):
        return self

    def isNull(self):
        return self.arr is None

    def width(self):
        return self.arr.shape[1]

    def height(self):
        return self.arr.shape[0]

    def __bool__(self):
        return self.arr is not None


class QListWidgetItem:
    def __init__(self, icon=None, text=''):
        self._text = str(text)
        self._data = {}

    def text(self):
        return self._text

    def setData(self, role, v):
        self._data[int(role)] = v

    def data(self, role):
        return self._data.get(int(role))

    def setIcon(self, *a):
        pass


class QListWidget:
    def __init__(self, *a, **k):
        self._items = []

    def clear(self):
        self._items = []

    def addItem(self, it):
        self._items.append(it)

    def count(self):
        return len(self._items)

    def item(self, i):
        return self._items[i] if 0 <= i < len(self._items) else None

    def currentRow(self):
        return -1

    def setCurrentRow(self, r):
        pass


class Spin:
    """Stands in for QSpinBox / QComboBox."""

    def __init__(self, v):
        self.v = v

    def value(self):
        return self.v

    def setValue(self, v):
        self.v = v

    def currentText(self):
        return self.v

    def setCurrentText(self, v):
        self.v = v

    def setRange(self, *a):
        pass


def _mod(name, **real):
    m = types.ModuleType(name)
    for k, v in real.items():
        setattr(m, k, v)
    m.__getattr__ = lambda n: Stub
    return m


def install():
    pkg = types.ModuleType('PyQt6')
    pkg.__path__ = []
    mods = {
        'QtCore': _mod('PyQt6.QtCore', QRect=QRect, QSize=QSize),
        'QtGui': _mod('PyQt6.QtGui', QImage=QImage, QPixmap=QPixmap),
        'QtWidgets': _mod('PyQt6.QtWidgets', QListWidget=QListWidget, QListWidgetItem=QListWidgetItem),
    }
    for k, m in mods.items():
        setattr(pkg, k, m)
        sys.modules['PyQt6.' + k] = m
    sys.modules['PyQt6'] = pkg
    return pkg
