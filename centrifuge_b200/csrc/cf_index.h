// cf_index.h -- host-side image of a Centrifuge `.1-.4.cf` index, flattened for upload.
//
// Format facts restated from the reference (paths relative to its tree):
//   .1.cf  bt2_io.h:138-526   header, plen, rstarts, ebwt sides, zOff, fchr, ftab, eftab
//   .2.cf  bt2_io.h:528-642   SA sample holding *sequence ids* (u16, or u32 iff nPat > 65535)
//   .3.cf  bt2_idx.h:623-707  uid->taxid, tree, names, sizes
//   .4.cf  bt2_idx.h:789-853  SA rows at genome starts -> sequence id
//   geometry EbwtParams::init bt2_idx.h:133-167
#ifndef CF_INDEX_H_
#define CF_INDEX_H_

#include <stdint.h>
#include <map>
#include <string>
#include <vector>

namespace cfb {

enum {
	RANK_UNKNOWN = 0, RANK_STRAIN, RANK_SPECIES, RANK_GENUS, RANK_FAMILY, RANK_ORDER, RANK_CLASS,
	RANK_PHYLUM, RANK_KINGDOM, RANK_DOMAIN, RANK_FORMA, RANK_INFRA_CLASS, RANK_INFRA_ORDER,
	RANK_PARV_ORDER, RANK_SUB_CLASS, RANK_SUB_FAMILY, RANK_SUB_GENUS, RANK_SUB_KINGDOM,
	RANK_SUB_ORDER, RANK_SUB_PHYLUM, RANK_SUB_SPECIES, RANK_SUB_TRIBE, RANK_SUPER_CLASS,
	RANK_SUPER_FAMILY, RANK_SUPER_KINGDOM, RANK_SUPER_ORDER, RANK_SUPER_PHYLUM, RANK_TRIBE,
	RANK_VARIETAS, RANK_LIFE, RANK_MAX
};
const char* rank_name(int rank);          // get_tax_rank_string taxonomy.h:207
int         rank_from_name(const char*);  // get_tax_rank_id     taxonomy.h:242
int         rank_to_slot(int rank);       // rank_to_pathID      taxonomy.h:68 (255 = none)

static const int kPathSlots = 10;         // TaxonomyPathTable::nranks

struct TaxNode { uint64_t taxid, parent; uint8_t rank, leaf; };

struct HostIndex {
	// ---- header / geometry
	uint64_t len = 0, bwt_len = 0, num_sides = 0, ftab_len = 0, eftab_len = 0, offs_len = 0;
	int32_t  line_rate = 0, off_rate = 0, ftab_chars = 0;
	uint64_t side_sz = 0, side_bwt_sz = 0, side_bwt_len = 0;
	uint64_t n_pat = 0;
	// ---- FM arrays
	std::vector<uint8_t>  sides;           // num_sides * side_sz, side 0 at offset 0
	uint64_t zoff = 0, fchr[5] = {0, 0, 0, 0, 0};
	std::vector<uint64_t> ftab, eftab;
	bool wide_sample = false;              // _offw
	std::vector<uint16_t> sample16; std::vector<uint32_t> sample32;
	// ---- genome-boundary rows (.4.cf), sorted by row
	std::vector<uint64_t> brow; std::vector<uint32_t> bseq;
	uint64_t last_boundary = 0; int bshift = 8; std::vector<uint32_t> bbits;
	// ---- taxonomy
	std::vector<std::string> seq_name; std::vector<uint64_t> seq_taxid;
	std::vector<TaxNode> nodes;            // sorted by taxid
	std::map<uint64_t, std::string> names; std::map<uint64_t, uint64_t> sizes;
	bool compressed = false;
	std::vector<int32_t>  seq_path;        // path id per sequence or -1 (taxid not in tree)
	std::vector<uint64_t> paths;           // n_paths * kPathSlots
	const TaxNode* find_node(uint64_t taxid) const;
	// load_cf_index(.., defer_bulk = true) leaves `sides` and the SA sample on disk and records where they are,
	// so that the device loader can stream them file -> pinned ring -> HBM without a host copy
	bool bulk_deferred = false; uint64_t sides_file_off = 0, sample_file_off = 0;
};

// Returns empty string on success, else the error message.
std::string load_cf_index(const std::string& basename, HostIndex& out, bool defer_bulk = false);

}  // namespace cfb
#endif
