// dy4_project.cpp — the reference's command-line boundary over the batched receiver, for ONE live stream:
//     rtl_sdr ... | dy4_project <mode> <mono|stereo> [blocks_per_call] | aplay ...
// Replaces the block loop of the reference's main(), src/project.cpp:289-318, and its stdin reader
// (readStdinBlockData, src/iofunc.cpp:113-120): whole blocks of interleaved uint8 I,Q come in on stdin, int16 PCM goes
// out on stdout, a trailing partial block is dropped and — as the reference does (project.cpp:293-296) — the process
// reports "End of input stream reached" and exits with status 1 at end of input.  Every sample is computed by
// libdy4b200.so's throughput tier (dy4_pipeline_process_host, n_streams = 1, carried state on the device) with
// DY4_FLAG_EXACT_AUDIO, so the bytes on stdout are the reference's.  blocks_per_call (default 1 = the reference's
// cadence, one 21-33 ms block per call) trades latency for fewer launches.  No CUDA code here: SURVEY.md §8f rank 2.
#include "../../include/dy4_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

int main(int argc, char** argv)
{
    if (argc < 3 || argc > 4) {
        std::fprintf(stderr, "Usage: %s <mode 0..3> <mono|stereo> [blocks_per_call]\n", argv[0]);
        return 1;
    }
    const int mode = std::atoi(argv[1]);
    const std::string channel = argv[2];
    const int per_call = argc == 4 ? std::atoi(argv[3]) : 1;
    if (mode < 0 || mode > 3) { std::fprintf(stderr, "Wrong mode: %d\n", mode); return 1; }          // project.cpp:160-162
    if (channel != "mono" && channel != "stereo") { std::fprintf(stderr, "Wrong parameter: %s, must be mono or stereo\n", channel.c_str()); return 1; }
    if (per_call < 1) { std::fprintf(stderr, "blocks_per_call must be >= 1\n"); return 1; }
    const int stereo = channel == "stereo";
    std::fprintf(stderr, "Operating in mode %d %s\n", mode, stereo ? "stereo" : "mono");              // project.cpp:176

    dy4_mode_params_t mp;
    dy4_pipeline_t* rx = nullptr;
    if (dy4_mode_params(mode, &mp) != DY4_OK ||
        dy4_pipeline_create(mode, stereo, 1, 0, DY4_FLAG_EXACT_AUDIO, &rx) != DY4_OK) {
        std::fprintf(stderr, "dy4_project: %s\n", dy4_last_error());
        return 2;
    }
    const size_t in_bytes = (size_t)per_call * mp.block_size;
    const size_t out_per_block = (size_t)mp.audio_per_block * (stereo ? 2 : 1);
    uint8_t* iq = (uint8_t*)dy4_pinned_alloc(in_bytes);
    int16_t* pcm = (int16_t*)dy4_pinned_alloc((size_t)per_call * out_per_block * sizeof(int16_t));
    if (!iq || !pcm) { std::fprintf(stderr, "dy4_project: %s\n", dy4_last_error()); return 2; }

    for (unsigned block_id = 0;; ) {
        const size_t got = std::fread(iq, 1, in_bytes, stdin);
        const int nb = (int)(got / mp.block_size);                                                     // whole blocks only
        if (nb > 0) {
            if (dy4_pipeline_process_host(rx, iq, in_bytes, nb, pcm, nullptr, 0) != DY4_OK) {
                std::fprintf(stderr, "dy4_project: %s\n", dy4_last_error());
                return 2;
            }
            std::fwrite(pcm, sizeof(int16_t), (size_t)nb * out_per_block, stdout);                    // project.cpp:317
            block_id += nb;
        }
        if (got < in_bytes) {
            std::fflush(stdout);
            std::fprintf(stderr, "End of input stream reached after %u blocks\n", block_id);          // project.cpp:293-296
            dy4_pinned_free(iq); dy4_pinned_free(pcm);
            dy4_pipeline_destroy(rx);
            return 1;
        }
    }
}
