// TEST INFRASTRUCTURE: the one websocketpp type the reference's client headers name.
#pragma once
#include <deque>
#include <map>
#include <memory>
namespace websocketpp { typedef std::weak_ptr<void> connection_hdl; }
