// cf_cli.cpp -- `centrifuge-class` drop-in executable: same name and argv conventions as the
// reference binary (centrifuge_main.cpp:42-68), so the stock Perl wrapper `centrifuge` can exec it.
#include "../../include/cfb200.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

int main(int argc, const char** argv) {
	if(argc > 2 && strcmp(argv[1], "-A") == 0) {   // batch mode: one argument line per job (centrifuge_main.cpp:43-64)
		std::ifstream in(argv[2]);
		std::string line; int last = 0;
		while(std::getline(in, line)) {
			if(line.empty() || line[0] == '#') continue;
			std::vector<std::string> toks; std::stringstream ss(line); std::string t;
			toks.push_back(argv[0]);
			while(ss >> t) toks.push_back(t);
			std::vector<const char*> av; for(size_t i = 0; i < toks.size(); i++) av.push_back(toks[i].c_str());
			last = cfb_run((int)av.size(), av.data());
			if(last) return last;
		}
		return last;
	}
	if(argc > 1 && strcmp(argv[1], "--promote") == 0) {   // centrifuge-promote <index> <tsv> <level> > out
		if(argc != 5) { std::cerr << "Usage: centrifuge-class --promote centrifuge_index_name centrifuge_output level > output" << std::endl; return 1; }
		return cfb_promote(argv[2], argv[3], argv[4], "-") == CFB_OK ? 0 : 1;
	}
	return cfb_run(argc, argv);
}
