// TEST INFRASTRUCTURE ONLY -- shim for spimap's Tree.h: the data members of spidir::Node / spidir::Tree that the
// reference's headers touch (src/MCMC.h:23-42,437,496,530-600).  No parsing, no arithmetic.
#pragma once
#include <string>
#include <vector>
namespace spidir {
struct Node {
    int name = 0;
    Node* parent = nullptr;
    Node** children = nullptr;
    int nchildren = 0;
    float dist = 0;
    std::string longname;
    bool isLeaf() const { return nchildren == 0; }
};
struct Tree {
    int nnodes = 0;
    Node* root = nullptr;
    std::vector<Node*> nodes;
    Tree(int n = 0) : nnodes(n) {}
    Node* getNode(int i) const { return nodes[(size_t)i]; }
};
}
