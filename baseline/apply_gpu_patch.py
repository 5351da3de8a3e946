#!/usr/bin/env python
"""Applies the GPU integration of INTEGRATION.md §2 to a COPY of the reference tree (never to /root/reference itself).

    python baseline/apply_gpu_patch.py <tree> <repo root>

The edits are insertions at textual anchors — only the added lines live here, no reference source is stored in this repo:

  CMakeLists.txt                               option DVBS2_LINK_CUDA: include dirs, libdvbs2_cuda.so, -DDVBS2_LINK_CUDA
  src/common/Factory/DVBS2/DVBS2.{hpp,cpp}     --dec-hw {CPU,GPU}, --gpu-sockets {FLOAT,SC16}, --gpu-ids, --sim-threads
                                               (next to --dec-simd, DVBS2.cpp:79-84 / :138-141)
  src/mains/TX_RX_BB/main.cpp                  the four decode tasks bound at :87-91 become one Decoder_chain_DVBS2_cuda task
                                               (or IQ_packer -> Decoder_chain_compact_DVBS2_cuda -> Bit_unpacker); thread count
                                               from --sim-threads instead of hardware_concurrency() (:19)
  src/mains/RX/main_sched.cpp (dvbs2_rx)       the same rebinding at :213-220; the scheduler then replicates the GPU task like
                                               any other (clone() = new handle, next GPU round-robin)
Every anchor must match exactly once; the script fails loudly otherwise (a changed reference needs a human look).
"""
import os
import sys


def edit(path, steps):
    s = open(path).read()
    for kind, anchor, text in steps:
        n = s.count(anchor)
        if n != 1:
            raise SystemExit("%s: anchor %r found %d times (expected 1)" % (path, anchor[:70], n))
        if kind == "after_line":      # insert after the line that contains the anchor
            i = s.index("\n", s.index(anchor)) + 1
            s = s[:i] + text + s[i:]
        elif kind == "before_line":   # insert before the line that contains the anchor
            i = s.rfind("\n", 0, s.index(anchor)) + 1
            s = s[:i] + text + s[i:]
        elif kind == "replace":
            s = s.replace(anchor, text)
        else:
            raise ValueError(kind)
    open(path, "w").write(s)


INCLUDES = '''#ifdef DVBS2_LINK_CUDA
#include "Decoder_chain_DVBS2_cuda.hpp" // dvbs2_b200: the fused GPU stage (libdvbs2_cuda)
#include "Adaptor_DVBS2_cuda.hpp"
#endif
'''

GPU_MODULES = '''
	// ---- dvbs2_b200: --dec-hw GPU builds the fused demap -> deinterleave -> LDPC -> BCH stage (one task, one handle per clone)
#ifdef DVBS2_LINK_CUDA
	const bool dec_gpu = params.dec_hw == "GPU";
	const bool gpu_sc16 = dec_gpu && params.gpu_sockets == "SC16";
	const float gpu_iq_scale = 1.f / 8192.f;
	uptr<dvbs2_b200::Decoder_chain_DVBS2_cuda> gpu_chain;
	uptr<dvbs2_b200::Decoder_chain_compact_DVBS2_cuda<int16_t>> gpu_cchain;
	uptr<dvbs2_b200::IQ_packer<int16_t>> gpu_packer;
	uptr<dvbs2_b200::Bit_unpacker> gpu_unpacker;
	if (dec_gpu && !gpu_sc16)
		gpu_chain.reset(new dvbs2_b200::Decoder_chain_DVBS2_cuda(params.modcod, params.ldpc_nite, params.ldpc_implem, 1.f, true,
		                                                         params.n_frames, params.gpu_id));
	if (gpu_sc16)
	{
		gpu_packer  .reset(new dvbs2_b200::IQ_packer<int16_t>(2 * params.N_xfec_frame, gpu_iq_scale, params.n_frames));
		gpu_cchain  .reset(new dvbs2_b200::Decoder_chain_compact_DVBS2_cuda<int16_t>(params.modcod, params.ldpc_nite,
		                   params.ldpc_implem, 1.f, true, params.n_frames, params.gpu_id, gpu_iq_scale, true));
		gpu_unpacker.reset(new dvbs2_b200::Bit_unpacker(params.K_bch, params.n_frames));
	}
#else
	const bool dec_gpu = false, gpu_sc16 = false;
	if (params.dec_hw == "GPU")
		throw spu::tools::runtime_error(__FILE__, __LINE__, __func__, "this binary was built without -DDVBS2_LINK_CUDA=ON.");
#endif
'''

# the float-socket rebinding (monitor_task = check_errors / check_errors2)
GPU_BIND = '''	}
#ifdef DVBS2_LINK_CUDA
	else if (!gpu_sc16)
	{
		(*gpu_chain   )[dvbs2_b200::chn::sck::demodulate_decode::CP  ] = (*estimator)[est::sck::estimate  ::SIG ];
		(*gpu_chain   )[dvbs2_b200::chn::sck::demodulate_decode::Y_N1] = (*framer   )[frm::sck::remove_plh::Y_N2];
		(*bb_scrambler)[scr::sck::descramble::Y_N1] = (*gpu_chain)[dvbs2_b200::chn::sck::demodulate_decode::V_K];
	}
	else
	{
		(*gpu_packer  )["pack::Y_N1"] = (*framer)[frm::sck::remove_plh::Y_N2];
		(*gpu_cchain  )[dvbs2_b200::chc::sck::demodulate_decode::CP  ] = (*estimator )[est::sck::estimate::SIG];
		(*gpu_cchain  )[dvbs2_b200::chc::sck::demodulate_decode::Y_N1] = (*gpu_packer)["pack::Y_N2"];
		(*gpu_unpacker)["unpack::X"] = (*gpu_cchain)[dvbs2_b200::chc::sck::demodulate_decode::V_K];
	}
#endif
'''


def monitor_v(task):
    return ('''	if (gpu_sc16) // the compact stage already descrambles: its unpacked payload goes straight to the monitor
#ifdef DVBS2_LINK_CUDA
		(*monitor)[mnt::sck::%s::V] = (*gpu_unpacker)["unpack::Y"];
#else
		;
#endif
	else
''' % task)


def main():
    tree, repo = sys.argv[1], os.path.abspath(sys.argv[2])
    j = lambda *p: os.path.join(tree, *p)  # noqa: E731

    with open(j("CMakeLists.txt"), "a") as f:
        f.write('''
# ---- dvbs2_b200: the GPU stage (libdvbs2_cuda + the header-only StreamPU/AFF3CT modules)
option(DVBS2_LINK_CUDA "Use libdvbs2_cuda for the demap/LDPC/BCH stage (--dec-hw GPU)" OFF)
if (DVBS2_LINK_CUDA)
    target_include_directories(dvbs2_common PUBLIC ${DVBS2_CUDA_ROOT}/include ${DVBS2_CUDA_ROOT}/dvbs2_b200/module)
    target_link_libraries     (dvbs2_common PUBLIC ${DVBS2_CUDA_ROOT}/dvbs2_b200/libdvbs2_cuda.so)
    target_compile_definitions(dvbs2_common PUBLIC DVBS2_LINK_CUDA)
    set(CMAKE_BUILD_RPATH "${DVBS2_CUDA_ROOT}/dvbs2_b200")
    foreach(t dvbs2_tx_rx_bb dvbs2_rx dvbs2_rx_sched)
        set_target_properties(${t} PROPERTIES BUILD_RPATH "${DVBS2_CUDA_ROOT}/dvbs2_b200;\\$ORIGIN/../../../../dvbs2_b200")
    endforeach()
endif()
''')

    edit(j("src/common/Factory/DVBS2/DVBS2.hpp"), [
        ("after_line", "std::string ldpc_simd;",
         "\tstd::string dec_hw;      // dvbs2_b200: CPU | GPU (the demap/LDPC/BCH stage on libdvbs2_cuda)\n"
         "\tstd::string gpu_sockets; // dvbs2_b200: FLOAT (AFF3CT socket types) | SC16 (sc16 symbols in, packed payload out)\n"
         "\tint         gpu_id;      // dvbs2_b200: CUDA device of the first clone, -1 = round-robin over all GPUs\n"
         "\tint         sim_threads; // dvbs2_b200: threads of the TX_RX_BB sequence, 0 = hardware_concurrency()\n"),
    ])
    edit(j("src/common/Factory/DVBS2/DVBS2.cpp"), [
        ("after_line", 'args.add({"dec-simd"}',
         '\targs.add({"dec-hw"},             cli::Text(cli::Including_set("CPU", "GPU")),                  "Where the demap/LDPC/BCH stage runs (GPU = libdvbs2_cuda)."       );\n'
         '\targs.add({"gpu-sockets"},        cli::Text(cli::Including_set("FLOAT", "SC16")),               "Socket formats of the GPU stage."                                );\n'
         '\targs.add({"gpu-id"},             cli::Integer(),                                               "CUDA device of the GPU stage (-1 = round-robin per clone)."      );\n'
         '\targs.add({"sim-threads"},        cli::Integer(cli::Positive()),                                "Threads of the simulation sequence (0 = all hardware threads)."  );\n'),
        ("after_line", 'ldpc_simd                = vals.exist({"dec-simd"}',
         '\tdec_hw                   = vals.exist({"dec-hw"}             ) ? vals.at      ({"dec-hw"}            ) : "CPU"       ;\n'
         '\tgpu_sockets              = vals.exist({"gpu-sockets"}        ) ? vals.at      ({"gpu-sockets"}       ) : "FLOAT"     ;\n'
         '\tgpu_id                   = vals.exist({"gpu-id"}             ) ? vals.to_int  ({"gpu-id"}            ) : -1          ;\n'
         '\tsim_threads              = vals.exist({"sim-threads"}        ) ? vals.to_int  ({"sim-threads"}       ) : 0           ;\n'),
    ])

    # ---------------------------------------------------------------- dvbs2_tx_rx_bb
    edit(j("src/mains/TX_RX_BB/main.cpp"), [
        ("after_line", '#include "Factory/DVBS2/DVBS2.hpp"', INCLUDES),
        ("after_line", "auto* LDPC_decoder = &LDPC_cdc->get_decoder_siho();", GPU_MODULES),
        ("before_line", "(*modem       )[mdm::sck::demodulate  ::CP  ]", "\tif (!dec_gpu)\n\t{\n"),
        ("after_line", "(*bb_scrambler)[scr::sck::descramble  ::Y_N1] = (*BCH_decoder )[", GPU_BIND),
        ("before_line", "(*monitor     )[mnt::sck::check_errors::V   ]", monitor_v("check_errors")),
        ("replace", "sequence_transmission((*source)[spu::module::src::tsk::generate], n_threads);",
         "sequence_transmission((*source)[spu::module::src::tsk::generate],\n"
         "\t                                             params.sim_threads > 0 ? (size_t)params.sim_threads : n_threads);"),
    ])

    # ---------------------------------------------------------------- dvbs2_rx (main_sched.cpp)
    edit(j("src/mains/RX/main_sched.cpp"), [
        ("after_line", '#include "Factory/DVBS2/DVBS2.hpp"', INCLUDES),
        ("after_line", "auto* LDPC_decoder = &LDPC_cdc->get_decoder_siho();", GPU_MODULES),
        ("before_line", "(*modem        )[             mdm::sck::demodulate   ::CP      ]", "\tif (!dec_gpu)\n\t{\n"),
        ("after_line", "(*bb_scrambler )[             scr::sck::descramble   ::Y_N1    ] = (*BCH_decoder  )[", GPU_BIND),
        ("before_line", "(*monitor      )[             mnt::sck::check_errors2::V       ]", monitor_v("check_errors2")),
        ("replace", "(*sink         )[spu::module::snk::sck::send         ::in_data ] = (*bb_scrambler )[             scr::sck::descramble   ::Y_N2    ];",
         "if (gpu_sc16)\n#ifdef DVBS2_LINK_CUDA\n\t\t(*sink)[spu::module::snk::sck::send::in_data] = (*gpu_unpacker)[\"unpack::Y\"];\n#else\n\t\t;\n#endif\n"
         "\telse\n\t\t(*sink         )[spu::module::snk::sck::send         ::in_data ] = (*bb_scrambler )[             scr::sck::descramble   ::Y_N2    ];"),
    ])
    print("patched", tree, "for", repo)


if __name__ == "__main__":
    main()
