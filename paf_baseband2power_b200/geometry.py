"""Block geometry of the BMF stream and the numbers derived from it.

Every constant is a reference constant: paf-baseband2power.conf:2-5 (NSAMP_DF,
NPOL_SAMP, NDIM_POL, NCHK_NIC), :9 (NDF), :24-25 (NCHAN, NBYTE); capture.h:27-32
(DF_SIZE 7232, DT_SIZE 7168, HDR_SIZE 64, TDF_SEC 1.08e-4); README.md:2
(27/32 us sampling, 1024x1024-sample integration).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Geometry:
    nchunk: int = 48          # NCHK_NIC
    nch_per_chunk: int = 7    # NCHAN / NCHK_NIC
    nsamp_df: int = 128       # NSAMP_DF
    npol_samp: int = 2        # NPOL_SAMP
    ndim_pol: int = 2         # NDIM_POL
    nbyte_in: int = 2         # 16-bit components
    ndf: int = 8192           # NDF, data frames per ring block
    nbyte_out: int = 4        # NBYTE, float32
    tsamp_us: float = 27.0 / 32.0

    @property
    def nchan(self) -> int:
        return self.nchunk * self.nch_per_chunk

    @property
    def pkt_bytes(self) -> int:
        """DT_SIZE: payload of one data frame of one chunk (7168)."""
        return self.nsamp_df * self.nch_per_chunk * self.npol_samp * self.ndim_pol * self.nbyte_in

    @property
    def frame_bytes(self) -> int:
        """One data frame of all chunks (344 064)."""
        return self.nchunk * self.pkt_bytes

    @property
    def block_bytes(self) -> int:
        """Input ring block: NDF*NCHK_NIC*7168 (paf-baseband2power.py:67)."""
        return self.ndf * self.frame_bytes

    @property
    def out_bytes(self) -> int:
        """Output ring block: NCHAN*NBYTE (paf-baseband2power.py:79)."""
        return self.nchan * self.nbyte_out

    @property
    def words_per_block(self) -> int:
        return self.block_bytes // 8

    @property
    def nsamp_integration(self) -> int:
        return self.ndf * self.nsamp_df

    @property
    def t_integration_s(self) -> float:
        """0.884736 s for the reference geometry (README.md:2)."""
        return self.nsamp_integration * self.tsamp_us * 1e-6

    @property
    def beam_rate_bytes_per_s(self) -> float:
        """Payload rate of one beam stream (3.1858 GB/s)."""
        return self.block_bytes / self.t_integration_s

    @property
    def algorithmic_bytes(self) -> int:
        """Bytes one beam-integration must move: block read + spectrum written."""
        return self.block_bytes + self.out_bytes


BMF = Geometry()
