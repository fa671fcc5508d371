// oracle/standin: DBoW2::BowVector (TEST INFRASTRUCTURE).  DBoW2 is a third-party dependency the reference expects
// under Thirdparty/DBoW2 and does not ship; restated from its published interface:
// a std::map<WordId, WordValue> with addWeight / addIfNotExist / normalize.
#pragma once
#include <cmath>
#include <map>
#include <vector>
using namespace std;   // the reference's include/Frame.h:37 writes `vector<size_t>` unqualified and relies on this
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
enum LNorm { L1, L2 };
enum WeightingType { TF_IDF, TF, IDF, BINARY };
enum ScoringType { L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT };
class BowVector : public std::map<WordId, WordValue> {
 public:
  void addWeight(WordId id, WordValue v) {
    auto vit = this->lower_bound(id);
    if (vit != this->end() && !(this->key_comp()(id, vit->first))) vit->second += v;
    else this->insert(vit, value_type(id, v));
  }
  void addIfNotExist(WordId id, WordValue v) {
    auto vit = this->lower_bound(id);
    if (vit == this->end() || (this->key_comp()(id, vit->first))) this->insert(vit, value_type(id, v));
  }
  void normalize(LNorm norm_type) {
    double norm = 0.0;
    if (norm_type == DBoW2::L1) { for (auto it = begin(); it != end(); ++it) norm += fabs(it->second); }
    else { for (auto it = begin(); it != end(); ++it) norm += it->second * it->second; norm = sqrt(norm); }
    if (norm > 0.0) for (auto it = begin(); it != end(); ++it) it->second /= norm;
  }
};
}  // namespace DBoW2
