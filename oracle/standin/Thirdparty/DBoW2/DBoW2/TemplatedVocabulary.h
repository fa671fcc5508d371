// oracle/standin: DBoW2::TemplatedVocabulary (TEST INFRASTRUCTURE).  Restated from DBoW2's published algorithm (the
// copy ORB-SLAM2 vendors under Thirdparty/DBoW2, absent from /root/reference): a k-ary tree of cluster centres, a
// feature descends by picking at every level the FIRST child at minimal distance (strict <), the leaf gives the word id
// and its weight, the ancestor `levelsup` levels above the leaves gives the FeatureVector node id; TF-IDF weighting adds
// the word weight per occurrence, L1 scoring normalises the vector by its L1 norm.
// Call sites pinned: src/Frame.cc:546-555, src/KeyFrame.cc:75-84 (transform(vDesc, mBowVec, mFeatVec, 4)).
#pragma once
#include <vector>
#include "BowVector.h"
#include "FeatureVector.h"
namespace DBoW2 {
template <class TDescriptor, class F> class TemplatedVocabulary {
 public:
  struct Node {
    NodeId id, parent;
    WordValue weight;
    std::vector<NodeId> children;
    TDescriptor descriptor;
    WordId word_id;
    Node() : id(0), parent(0), weight(0), word_id(0) {}
    bool isLeaf() const { return children.empty(); }
  };
  TemplatedVocabulary(int k = 10, int L = 5, WeightingType weighting = TF_IDF, ScoringType scoring = L1_NORM)
      : m_k(k), m_L(L), m_weighting(weighting), m_scoring(scoring) {}
  bool empty() const { return m_words.empty(); }
  unsigned int size() const { return (unsigned int)m_words.size(); }
  // test-side construction: nodes in id order, node 0 = root; parent[i] < i; leaves become words in id order
  void build(int k, int L, const std::vector<NodeId>& parent, const std::vector<TDescriptor>& desc,
             const std::vector<WordValue>& weight) {
    m_k = k; m_L = L;
    m_nodes.assign(parent.size(), Node());
    m_words.clear();
    for (size_t i = 0; i < parent.size(); ++i) {
      m_nodes[i].id = (NodeId)i; m_nodes[i].parent = parent[i]; m_nodes[i].weight = weight[i]; m_nodes[i].descriptor = desc[i];
      if (i) m_nodes[parent[i]].children.push_back((NodeId)i);
    }
    for (size_t i = 0; i < m_nodes.size(); ++i)
      if (i && m_nodes[i].isLeaf()) { m_nodes[i].word_id = (WordId)m_words.size(); m_words.push_back(&m_nodes[i]); }
  }
  virtual ~TemplatedVocabulary() {}
  virtual void transform(const std::vector<TDescriptor>& features, BowVector& v, FeatureVector& fv, int levelsup) const {
    v.clear(); fv.clear();
    if (empty()) return;
    const bool must = (m_scoring == L1_NORM || m_scoring == L2_NORM);
    const LNorm norm = (m_scoring == L2_NORM) ? L2 : L1;
    unsigned int i_feature = 0;
    for (auto fit = features.begin(); fit < features.end(); ++fit, ++i_feature) {
      WordId id; NodeId nid = 0; WordValue w;
      transform(*fit, id, w, &nid, levelsup);
      if (m_weighting == TF_IDF || m_weighting == TF) { if (w > 0) { v.addWeight(id, w); fv.addFeature(nid, i_feature); } }
      else { if (w > 0) { v.addIfNotExist(id, w); fv.addFeature(nid, i_feature); } }
    }
    if ((m_weighting == TF_IDF || m_weighting == TF) && !v.empty() && !must) {
      const double nd = (double)v.size();
      for (auto vit = v.begin(); vit != v.end(); ++vit) vit->second /= nd;
    }
    if (must) v.normalize(norm);
  }
  void transform(const TDescriptor& feature, WordId& word_id, WordValue& weight, NodeId* nid, int levelsup) const {
    const int nid_level = m_L - levelsup;
    if (nid_level <= 0 && nid != NULL) *nid = 0;
    NodeId final_id = 0;
    int current_level = 0;
    do {
      ++current_level;
      const std::vector<NodeId>& nodes = m_nodes[final_id].children;
      final_id = nodes[0];
      double best_d = F::distance(feature, m_nodes[final_id].descriptor);
      for (auto nit = nodes.begin() + 1; nit != nodes.end(); ++nit) {
        NodeId id = *nit;
        double d = F::distance(feature, m_nodes[id].descriptor);
        if (d < best_d) { best_d = d; final_id = id; }
      }
      if (nid != NULL && current_level == nid_level) *nid = final_id;
    } while (!m_nodes[final_id].isLeaf());
    word_id = m_nodes[final_id].word_id;
    weight = m_nodes[final_id].weight;
  }
  double score(const BowVector& v1, const BowVector& v2) const {   // L1 scoring: 1 - 0.5 * | v1 - v2 |_1
    double score = 0;
    auto v1_it = v1.begin(), v2_it = v2.begin();
    while (v1_it != v1.end() && v2_it != v2.end()) {
      if (v1_it->first == v2_it->first) {
        score += fabs(v1_it->second - v2_it->second) - fabs(v1_it->second) - fabs(v2_it->second);
        ++v1_it; ++v2_it;
      } else if (v1_it->first < v2_it->first) v1_it = v1.lower_bound(v2_it->first);
      else v2_it = v2.lower_bound(v1_it->first);
    }
    return -score / 2.0;
  }
protected:
  int m_k, m_L;
  WeightingType m_weighting;
  ScoringType m_scoring;
  std::vector<Node> m_nodes;
  std::vector<Node*> m_words;
};
}  // namespace DBoW2
