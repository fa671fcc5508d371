// oracle/standin: DBoW2::FeatureVector (TEST INFRASTRUCTURE): std::map<NodeId, std::vector<unsigned int>> + addFeature.
#pragma once
#include "BowVector.h"
namespace DBoW2 {
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {
 public:
  void addFeature(NodeId id, unsigned int i_feature) {
    auto vit = this->lower_bound(id);
    if (vit != this->end() && vit->first == id) vit->second.push_back(i_feature);
    else { vit = this->insert(vit, value_type(id, std::vector<unsigned int>())); vit->second.push_back(i_feature); }
  }
};
}  // namespace DBoW2
