"""rtlsdr_wsprd_b200 -- B200 (sm_100a) batch WSPR decoder behind the entry points of Guenael/rtlsdr-wsprd.

The product is ``libwsprd_b200.so`` (CUDA kernels + C ABI, see ``include/wspr_b200.h``); this package is the thin
Python host side over that C ABI: struct mirrors of ``wsprd/wsprd.h:44-74``, ``wspr_decode`` with the reference's
argument meaning, and the batch / device-resident context a throughput caller uses.  There is no CPU fallback:
every compute call raises ``WsprCudaError`` when the library or a CUDA device is missing.
"""
from .wsprd import (DecoderOptions, DecoderResults, RESULT_DTYPE, CAND_DTYPE, MAX_UNIQUES, NSAMP, WsprCudaError,
                    default_options, library, library_path, wspr_decode, decode_batch, BatchDecoder, PipelinedDecoder, decimate_batch,
                    decimate_device, FrontEnd, fano_batch, fano_pool_stats, spot_line, print_spots_lines, wsprnet_urls, read_iq_file, read_c2_file, write_iq_file, write_c2_file, load_capture_files, normalise_half,
                    kernel_launches, build_library)

__all__ = ["DecoderOptions", "DecoderResults", "RESULT_DTYPE", "CAND_DTYPE", "MAX_UNIQUES", "NSAMP", "WsprCudaError",
           "default_options", "library", "library_path", "wspr_decode", "decode_batch", "BatchDecoder", "PipelinedDecoder",
           "decimate_batch", "decimate_device", "FrontEnd", "fano_batch", "fano_pool_stats", "spot_line", "print_spots_lines", "wsprnet_urls", "read_iq_file", "read_c2_file", "write_iq_file", "write_c2_file", "load_capture_files",
           "normalise_half", "kernel_launches", "build_library"]
