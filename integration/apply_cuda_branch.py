#!/usr/bin/env python3
"""Writes patched copies of the reference's factory sources with the "cuda" branches of INTEGRATION.md section 2.

    apply_cuda_branch.py <reference root> <output dir>

The reference tree is read-only and none of its code is kept in this repository: the script only knows the function
signatures after which a branch is inserted (anchors), reads the file where it lies and writes the patched copy into
the build directory. What it inserts is exactly what a maintainer would add by hand.
"""
import sys
from pathlib import Path

INCLUDE = '#include "pusch_dec_cuda_adapters.h" // B200 "cuda" variants\n'

SHARED_CONTEXT = '''
namespace {
/// One GPU context shared by every "cuda" object of this process (nullptr without a usable GPU: no software fallback).
std::shared_ptr<srsran::cuda::context> cuda_context()
{
  static std::shared_ptr<srsran::cuda::context> ctx = srsran::cuda::context::create({});
  return ctx;
}
} // namespace
'''

# file -> list of (anchor line prefix, lines inserted after the opening brace that follows the anchor)
PATCHES = {
    "lib/phy/upper/channel_coding/channel_coding_factories.cpp": [
        ("std::shared_ptr<ldpc_decoder_factory> srsran::create_ldpc_decoder_factory_sw(",
         '  if (dec_type == "cuda") {\n    return srsran::cuda::create_ldpc_decoder_factory_cuda(cuda_context());\n  }\n'),
        ("std::shared_ptr<ldpc_encoder_factory> srsran::create_ldpc_encoder_factory_sw(",
         '  if (enc_type == "cuda") {\n    return srsran::cuda::create_ldpc_encoder_factory_cuda(cuda_context());\n  }\n'),
        ("srsran::create_ldpc_rate_dematcher_factory_sw(",
         '  if (dematcher_type == "cuda") {\n'
         '    return srsran::cuda::create_ldpc_rate_dematcher_factory_cuda(cuda_context());\n  }\n'),
        ("std::shared_ptr<crc_calculator_factory> srsran::create_crc_calculator_factory_sw(",
         '  if (type == "cuda") {\n    return srsran::cuda::create_crc_calculator_factory_cuda(cuda_context());\n  }\n'),
    ],
}


# The hal registry (create_hw_accelerator_pusch_dec_factory keyed by acc_type, "acc100" today): its whole body sits under
# ENABLE_PUSCH_HWACC, i.e. needs DPDK; the "cuda" accelerator does not, so its branch goes in front of the #ifdef.
HAL_FILE = "lib/hal/phy/upper/channel_processors/pusch/hw_accelerator_factories.cpp"
HAL_ANCHOR = "srsran::hal::create_hw_accelerator_pusch_dec_factory("
HAL_BODY = ('  if (accelerator_config.acc_type == "cuda") {\n'
            '    // B200: batched LDPC decoding + rate dematching + HARQ in device memory (external soft bits), no DPDK.\n'
            '    return srsran::cuda::create_hw_accelerator_pusch_dec_factory_cuda(cuda_context());\n  }\n')
PATCHES[HAL_FILE] = [(HAL_ANCHOR, HAL_BODY)]


# The reference's pusch_decoder_hwacc_benchmark knows one accelerator name ("acc100"); a maintainer adding the CUDA
# accelerator adds its name next to it, going through the SAME hal registry and pusch_decoder_hw_impl as ACC100 does.
HWACC_BENCH = "tests/benchmarks/phy/upper/channel_processors/pusch/pusch_decoder_hwacc_benchmark.cpp"
HWACC_ANCHOR = "static std::shared_ptr<pusch_decoder_factory> create_pusch_decoder_factory(std::string decoder_type)"
HWACC_BODY = ('  if (decoder_type == "cuda") {\n'
              '    hal::hw_accelerator_pusch_dec_configuration hw_decoder_config;\n'
              '    hw_decoder_config.acc_type       = "cuda";\n'
              '    hw_decoder_config.ext_softbuffer = true;\n'
              '    pusch_decoder_factory_hw_configuration decoder_hw_factory_config;\n'
              '    decoder_hw_factory_config.segmenter_factory  = create_ldpc_segmenter_rx_factory_sw();\n'
              '    decoder_hw_factory_config.crc_factory        = create_crc_calculator_factory_sw("auto");\n'
              '    decoder_hw_factory_config.hw_decoder_factory = hal::create_hw_accelerator_pusch_dec_factory(hw_decoder_config);\n'
              '    TESTASSERT(decoder_hw_factory_config.hw_decoder_factory, "No CUDA accelerator (no GPU?).");\n'
              '    return create_pusch_decoder_factory_hw(decoder_hw_factory_config);\n  }\n')
NO_CONTEXT = {HWACC_BENCH}  # files that do not need the shared context / adapter include
PATCHES[HWACC_BENCH] = [(HWACC_ANCHOR, HWACC_BODY)]


def patch(text, rules, with_context=True):
    lines = text.splitlines(keepends=True)
    out, pending, done = [], None, 0
    first_include = min(i for i, l in enumerate(lines) if l.startswith("#include"))
    for i, line in enumerate(lines):
        out.append(line)
        if i == first_include and with_context:
            out.append(INCLUDE)
            out.append(SHARED_CONTEXT)
        for anchor, body in rules:
            if line.startswith(anchor):
                pending = body
        if pending is not None and line.strip() == "{":
            out.append(pending)
            pending, done = None, done + 1
    if done != len(rules):
        raise SystemExit("anchors not found: %d of %d applied" % (done, len(rules)))
    return "".join(out)


def main():
    ref, dst = Path(sys.argv[1]), Path(sys.argv[2])
    dst.mkdir(parents=True, exist_ok=True)
    for rel, rules in PATCHES.items():
        target = dst / rel.replace("/", "__")
        target.write_text(patch((ref / rel).read_text(), rules, rel not in NO_CONTEXT))
        print("patched", rel, "->", target)


if __name__ == "__main__":
    main()
