"""MatrixMarket 'array real general' reader/writer (oracle, test infrastructure).

Follows votca::tools::EigenIO_MatrixMarket (tools/src/libtools/
eigenio_matrixmarket.cc:30-75): header line, optional % comments, "rows cols",
then rows*cols values listed column-major.
"""
import numpy as np


def read_matrix(path):
    with open(path) as fh:
        lines = [ln.strip() for ln in fh if ln.strip() and not ln.startswith("%")]
    rows, cols = (int(x) for x in lines[0].split()[:2])
    vals = np.array([float(x) for ln in lines[1:] for x in ln.split()], dtype=np.float64)
    if vals.size != rows * cols:
        raise ValueError(f"{path}: expected {rows * cols} values, got {vals.size}")
    return vals.reshape((rows, cols), order="F")


def read_vector(path):
    return read_matrix(path).reshape(-1, order="F")


def write_matrix(path, mat):
    mat = np.atleast_2d(np.asarray(mat, dtype=np.float64))
    with open(path, "w") as fh:
        fh.write("%%MatrixMarket matrix array real general\n")
        fh.write(f"{mat.shape[0]} {mat.shape[1]}\n")
        for v in mat.reshape(-1, order="F"):
            fh.write(f"{v:.17g}\n")
