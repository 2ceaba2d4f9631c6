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


# The gNB itself: the upper PHY builds its PUSCH decoder factory from the configured LDPC decoder type.
# With "cuda" it takes the reference's hardware front end (pusch_decoder_hw_impl) on the CUDA accelerator from the hal
# registry - nothing else of the upper PHY changes - and the rx buffer pool keeps its soft bits in the accelerator
# (external_soft_bits), which is what du_low_config_translator.cpp has to request.
UPPER_PHY = "lib/phy/upper/upper_phy_factories.cpp"
UPPER_PHY_LINE = "  pusch_config.decoder_factory                      = create_pusch_decoder_factory_sw(decoder_config);"
UPPER_PHY_NEW = ('  if (config.ldpc_decoder_type == "cuda") {\n'
                 '    hal::hw_accelerator_pusch_dec_configuration hw_cfg;\n'
                 '    hw_cfg.acc_type       = "cuda";\n'
                 '    hw_cfg.ext_softbuffer = true;\n'
                 '    pusch_decoder_factory_hw_configuration hw_decoder_config;\n'
                 '    hw_decoder_config.segmenter_factory  = decoder_config.segmenter_factory;\n'
                 '    hw_decoder_config.crc_factory        = crc_calc_factory;\n'
                 '    hw_decoder_config.hw_decoder_factory = hal::create_hw_accelerator_pusch_dec_factory(hw_cfg);\n'
                 '    report_fatal_error_if_not(hw_decoder_config.hw_decoder_factory, "No CUDA PUSCH decoder accelerator.");\n'
                 '    pusch_config.decoder_factory = create_pusch_decoder_factory_hw(hw_decoder_config);\n'
                 '  } else {\n'
                 '    pusch_config.decoder_factory = create_pusch_decoder_factory_sw(decoder_config);\n'
                 '  }\n')
UPPER_PHY_INCLUDE = '#include "srsran/hal/phy/upper/channel_processors/pusch/hw_accelerator_factories.h"\n'
DU_LOW = "apps/units/flexible_du/du_low/du_low_config_translator.cpp"
DU_LOW_LINES = [
    ("    upper_phy_cell.rx_buffer_config.external_soft_bits   = false;",
     '    // SRSRAN_LDPC_DECODER_TYPE=cuda selects the B200 PUSCH decoder (until expert_phy grows the three type fields)\n'
     '    const char* pdc_type = std::getenv("SRSRAN_LDPC_DECODER_TYPE");\n'
     '    const bool  pdc_cuda = pdc_type != nullptr && std::string(pdc_type) == "cuda";\n'
     '    upper_phy_cell.rx_buffer_config.external_soft_bits   = pdc_cuda;\n'),
    ('    upper_phy_cell.ldpc_rate_dematcher_type              = "auto";',
     '    upper_phy_cell.ldpc_rate_dematcher_type              = pdc_cuda ? "cuda" : "auto";\n'),
    ('    upper_phy_cell.ldpc_decoder_type                     = "auto";',
     '    upper_phy_cell.ldpc_decoder_type                     = pdc_cuda ? "cuda" : "auto";\n'),
]


def replace_lines(text, rules, extra_include=None):
    lines = text.splitlines(keepends=True)
    out, done = [], 0
    first_include = min(i for i, l in enumerate(lines) if l.startswith("#include"))
    for i, line in enumerate(lines):
        hit = [new for old, new in rules if line.rstrip("\n") == old]
        if hit:
            out.append(hit[0])
            done += 1
        else:
            out.append(line)
        if i == first_include and extra_include:
            out.append(extra_include)
    if done != len(rules):
        raise SystemExit("lines not found: %d of %d replaced" % (done, len(rules)))
    return "".join(out)


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
    for rel, rules, inc in ((UPPER_PHY, [(UPPER_PHY_LINE, UPPER_PHY_NEW)], UPPER_PHY_INCLUDE),
                            (DU_LOW, DU_LOW_LINES, "#include <cstdlib>\n#include <string>\n")):
        target = dst / rel.replace("/", "__")
        target.write_text(replace_lines((ref / rel).read_text(), rules, inc))
        print("patched", rel, "->", target)


if __name__ == "__main__":
    main()
