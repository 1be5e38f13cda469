// Peer-memory transport of the row-strip sharded forward: see peer.cu.
#pragma once
#include "common.cuh"

#include <algorithm>
#include <cstring>

namespace cidnet {

static constexpr int kPeerMaxRanks = 8;
static constexpr int kPeerMaxJobs = 16;
static constexpr int64_t kPeerHdrBytes = 4096;      // synchronisation header at the start of every rank's workspace

struct PeerHdr {
    uint32_t seq;                    // number of the last exchange this rank has completed
    uint32_t arrive;                 // CTA arrival counter of the exchange in flight (last-CTA detection)
    uint32_t error;                  // set when a handshake timed out
    uint32_t pad;
    uint32_t ready[kPeerMaxRanks];   // ready[r]: written by rank r -- "my data of exchange v is complete"
    uint32_t ack[kPeerMaxRanks];     // ack[r]:   written by rank r -- "I have read your data of exchange v"
};

struct PeerSync {
    PeerHdr* me;
    PeerHdr* partner[kPeerMaxRanks];     // the partners' headers (peer mappings)
    int partner_rank[kPeerMaxRanks];
    int npartners, rank;
};

struct PeerCopyJob { void* dst; const void* src; long long bytes; };   // dst local, src in a partner's workspace
struct PeerHaloArgs { PeerSync sync; PeerCopyJob job[kPeerMaxJobs]; int njobs; };
int launch_peer_halo(const PeerHaloArgs& a, cudaStream_t stream);

struct PeerReduceArgs { PeerSync sync; const float* src[kPeerMaxRanks]; float* dst; int count, nranks; };
int launch_peer_allreduce(const PeerReduceArgs& a, cudaStream_t stream);

}  // namespace cidnet
