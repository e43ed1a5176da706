// oracle/ref_mock/DataStore.h — TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's include/DataStore.h (dataset loaders,
// boost::filesystem ...): include/Event/EventData.h needs only the SharedQueue<T> base of its EventQueue (EventData.h:127-137),
// which the event-frame path (EventConversion.cc) never touches.
#pragma once
#include <iostream>
#include <vector>
#include <opencv2/core/core.hpp>

namespace EORB_SLAM {
template <class T> class SharedQueue {
public:
    virtual ~SharedQueue() = default;
    void fillBuffer(const std::vector<T>&) {}
    unsigned long consumeBegin(unsigned long, std::vector<T>&) { return 0; }
};
}  // namespace EORB_SLAM
