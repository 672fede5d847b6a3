"""fastqandfurious_b200 -- B200-native FASTQ-buffer parser behind the fastq-and-furious plugin API.

``from fastqandfurious_b200 import readfastq_iter, entryfunc, entrypos`` replaces
``from fastqandfurious import ...`` / ``from fastqandfurious._fastqandfurious import entrypos``.
The CUDA extension (libfqb200.so) is mandatory: there is no CPU fallback."""
from . import _lib, consume, device  # noqa: F401
from ._lib import (COMPLETE, INVALID, MISSING_QUAL_BEGIN, MISSING_QUAL_END, MISSING_QUALHEADER_END,  # noqa: F401
                   MISSING_SEQ_BEG, MISSING_SEQ_END, MISSING_SEQHEADER_BEGIN, MISSING_SEQHEADER_END, POS_HEAD_BEG,
                   POS_HEAD_END, POS_QUAL_BEG, POS_QUAL_END, POS_SEQ_BEG, POS_SEQ_END)
from .api import (FORMAT_OPENERS, DeviceEntryPos, DeviceEntryPosFasta, Entry, arrayadd_b, arrayadd_q,  # noqa: F401
                  automagic_open, entryfunc, entryfunc_abspos, entryfunc_fasta, entryfunc_namedtuple, entrypos,
                  entrypos_fasta, read, readfastq_iter, readfastq_table)
from .consume import (field_lengths, field_sums, gather_fields, pack_2bit, read_index, select_by_length,  # noqa: F401
                      write_index)
from .device import FastaResult, ParseResult, parse_buffer, parse_fasta_buffer, synth_fixed  # noqa: F401


class _CExtNamespace:
    """Stand-in for the reference's ``fastqandfurious._fastqandfurious`` module surface
    (src/_fastqandfurious.c:220-265)."""
    entrypos = staticmethod(entrypos)
    arrayadd_b = staticmethod(arrayadd_b)
    arrayadd_q = staticmethod(arrayadd_q)
    INVALID = INVALID
    POS_HEAD_BEG, POS_HEAD_END, POS_SEQ_BEG, POS_SEQ_END = POS_HEAD_BEG, POS_HEAD_END, POS_SEQ_BEG, POS_SEQ_END
    POS_QUAL_BEG, POS_QUAL_END = POS_QUAL_BEG, POS_QUAL_END
    COMPLETE, MISSING_QUALHEADER_END = COMPLETE, MISSING_QUALHEADER_END


_fastqandfurious = _CExtNamespace()
__version__ = '0.1.0'
