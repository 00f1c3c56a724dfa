import enum


class MatrixResizeType(enum.Enum):
    SET_ZERO = 0
    UNDEFINED = 1
    COPY_DATA = 2


class MatrixTransposeType(enum.Enum):
    NO_TRANS = 111
    TRANS = 112
