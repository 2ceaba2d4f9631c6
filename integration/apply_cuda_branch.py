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


def patch(text, rules):
    lines = text.splitlines(keepends=True)
    out, pending, done = [], None, 0
    first_include = min(i for i, l in enumerate(lines) if l.startswith("#include"))
    for i, line in enumerate(lines):
        out.append(line)
        if i == first_include:
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
        target.write_text(patch((ref / rel).read_text(), rules))
        print("patched", rel, "->", target)


if __name__ == "__main__":
    main()
