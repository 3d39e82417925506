"""kf2vecfsw_b200 -- B200-native k-mer frequency step for kf2vec (drop-in for
``kf2vec.main.get_frequencies``, reference ``kf2vec/main.py:250-373``).

Host code is Python over a C-ABI shared library (``libkfcount.so``, see ``include/kfcount.h``) whose
hot path is hand-written CUDA for sm_100a.  There is no CPU fallback: importing works anywhere, but
every compute call raises if the library is not built or no B200 is visible.
"""
from .engine import (  # noqa: F401
    KfError,
    count_buffers,
    count_device,
    count_files,
    count_windows,
    parse_kf,
    last_file_status,
    DeviceArena,
    format_row,
    init,
    shutdown,
    last_count_kernel_ms,
    last_launch_count,
    lib_path,
    vocab,
    vocab_codes,
    vocab_size,
    write_kf,
    write_kf_rows,
    linearise_fasta,
)
from .frequencies import get_frequencies, frequency_matrix  # noqa: F401
from .chunks import get_chunks  # noqa: F401
from .kmers import get_kmers  # noqa: F401
from .loader import load_kf_dir, load_kf_files, read_chunk_kf, read_kf  # noqa: F401

__version__ = "0.1.0"
