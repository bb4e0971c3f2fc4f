// TEST INFRASTRUCTURE ONLY.
//
// Link-time tap on muscle::GlobalAlign (MU/glbalign.cpp:69-81) for minting DP fixtures from the REAL pipeline: the object file
// of glbalign.cpp is copied with its GlobalAlign symbol renamed (objcopy, oracle/Makefile.ref) and this file supplies
// GlobalAlign: it forwards to the renamed original and appends, for every call, the two profiles as letter strings (when every
// column is a single ungapped ACGT letter, the case mcu_nw_batch covers; otherwise "-" and the column counts) and the path the
// reference returned, to the file named by $MAUVE_DP_TRACE.  No source of the reference is touched.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "libMUSCLE/muscle.h"
#include "libMUSCLE/profile.h"
#include "libMUSCLE/pwpath.h"

namespace muscle {

SCORE GlobalAlign_reference(const ProfPos* PA, unsigned uLengthA, const ProfPos* PB, unsigned uLengthB, PWPath& Path)
    asm("_ZN6muscle21GlobalAlign_referenceEPKNS_7ProfPosEjS2_jRNS_6PWPathE");

static bool letters(const ProfPos* P, unsigned n, std::string& out)
{
	out.resize(n);
	for (unsigned i = 0; i < n; ++i) {
		const unsigned u = P[i].m_uSortOrder[0];
		if (P[i].m_bAllGaps || u >= 4 || P[i].m_fcCounts[u] != 1.0f) return false;
		out[i] = "ACGT"[u];
	}
	return n > 0;
}

SCORE GlobalAlign(const ProfPos* PA, unsigned uLengthA, const ProfPos* PB, unsigned uLengthB, PWPath& Path)
{
	static FILE* f = NULL;
	static bool tried = false;
	if (!tried) {
		tried = true;
		const char* p = getenv("MAUVE_DP_TRACE");
		if (p) f = fopen(p, "w");
	}
	std::string a, b;
	const bool ok = f && letters(PA, uLengthA, a) && letters(PB, uLengthB, b);   // read before SetTermGaps touches the profiles
	const SCORE s = GlobalAlign_reference(PA, uLengthA, PB, uLengthB, Path);
	if (f) {
		if (ok) {
			std::string path(Path.GetEdgeCount(), '?');
			for (unsigned e = 0; e < Path.GetEdgeCount(); ++e) path[e] = Path.GetEdge(e).cType;
			fprintf(f, "%s %s %s\n", a.c_str(), b.c_str(), path.c_str());
		} else
			fprintf(f, "- %u %u\n", uLengthA, uLengthB);
		fflush(f);
	}
	return s;
}

}  // namespace muscle
